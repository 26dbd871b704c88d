#!/bin/bash
# LBVH half of the round-1 (e) evidence run (see profile_r01e.sh): GPU tests, config-5 bench line, --set full captures.
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_r01e.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_r01e.log; tail -3 $O/pytest_r01e.log
python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/bench_r01e_cfg5.json 2> $O/bench_r01e_cfg5.err; echo "bench cfg5 rc=$?"
python bench.py --workload cfg2 --accel lbvh --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/bench_r01e_cfg2_lbvh.json 2> $O/bench_r01e_cfg2_lbvh.err; echo "bench cfg2 lbvh rc=$?"
B=raytracingincuda_b200/bin/b200-raytrace
CLI="$B --scene_id 1 --width 1920 --height 1080 --samples 16 --bounces 25 --no-ppm --stats"
$CLI --accel lbvh > $O/plain_pb_l1.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:trace_kernel_pb -o $O/prof_r01e_pb_lbvh_s1 -f $CLI --accel lbvh > $O/ncu_pb_l1.log 2>&1
echo "ncu lbvh s1 rc=$?"
$CLI --scaled_half 158 > $O/plain_pb_l2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:trace_kernel_pb -o $O/prof_r01e_pb_lbvh_100k -f $CLI --scaled_half 158 > $O/ncu_pb_l2.log 2>&1
echo "ncu lbvh 100k rc=$?"
