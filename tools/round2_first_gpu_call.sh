#!/bin/bash
# First GPU call of the next round: the uniform grid (RT_ACCEL_GRID, csrc/rt_grid.cuh) with the per-step inflation has not
# run on hardware yet.  Gated parity tests, timings next to the LBVH, bench lines and one ncu capture of the grid kernel.
#   gpurun --timeout 600 -- 'bash tools/round2_first_gpu_call.sh'
set -u
O=gpurun_out
export RT_ENABLE_GRID=1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grid" 2>&1 | tail -5 | tee $O/r02_grid_tests.log
python tools/time_grid.py 2>&1 | tee $O/r02_grid_timing.log
python bench.py --workload cfg2 --accel grid --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-lbvh-extra > $O/r02_bench_cfg2_grid.json 2> $O/r02_bench_cfg2_grid.err
python bench.py --workload cfg5 --accel grid --steps 2 --warmup 2 --no-cpu-baseline --no-ref-gpu --no-lbvh-extra > $O/r02_bench_cfg5_grid.json 2> $O/r02_bench_cfg5_grid.err
B=raytracingincuda_b200/bin/b200-raytrace
CLI="$B --scene_id 1 --width 1920 --height 1080 --samples 16 --bounces 25 --no-ppm --stats --accel grid"
$CLI > $O/plain_grid.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:trace_kernel_pb -o $O/prof_r02_grid_s1 -f $CLI > $O/ncu_grid.log 2>&1
echo "ncu rc=$?"
