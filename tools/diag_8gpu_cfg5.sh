#!/bin/bash
# Why do the ranks of an 8-GPU box take 158-249 ms for the same config-5 share?  The same render on every GPU, alone and then
# all at once, LBVH and grid, with clocks.   usage: gpurun --gpus 8 -- 'bash tools/diag_8gpu_cfg5.sh'
B=raytracingincuda_b200/bin/b200-raytrace
ARGS="--scene_id 1 --scaled_half 158 --width 3840 --height 2160 --samples 32 --bounces 50 --no-ppm"
for ACC in lbvh grid; do
  echo "== alone, --accel $ACC"
  for G in 0 1 2 3 4 5 6 7; do
    CUDA_VISIBLE_DEVICES=$G $B $ARGS --accel $ACC > /dev/null      # warm (context, build)
    echo "gpu $G: $(CUDA_VISIBLE_DEVICES=$G $B $ARGS --accel $ACC | cut -d, -f1)"
  done
  echo "== all eight at once, --accel $ACC"
  for G in 0 1 2 3 4 5 6 7; do
    (CUDA_VISIBLE_DEVICES=$G $B $ARGS --accel $ACC --samples 128 | cut -d, -f1 | sed "s/^/gpu $G (128 spp): /") &
  done
  sleep 0.6
  nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv,noheader
  wait
done
