#!/bin/bash
# Multi-GPU evidence on N B200 of one box: the CLI's two splits (P2P + host paths) against the single-GPU PPM, and the bench line
# with both partitionings and config 5.   usage: gpurun --gpus N -- 'bash tools/r02_multi_gpu.sh N [tag]'
set -u
N=${1:-2}; T=${2:-r02}
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cli_multi_gpu or spp_split or row_split" 2>&1 | tail -3 | tee $O/${T}_multi${N}_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 \
    > $O/${T}_bench_${N}gpu.json 2> $O/${T}_bench_${N}gpu.err
echo "bench $N rc=$?"; tail -c 600 $O/${T}_bench_${N}gpu.err
B=raytracingincuda_b200/bin/b200-raytrace
for SP in rows spp; do
  $B --scene_id 1 --width 3840 --height 2160 --samples 1000 --bounces 50 --no-ppm --stats --gpus $N --split $SP 2>&1 | tee $O/${T}_cli_${N}gpu_$SP.log
done
