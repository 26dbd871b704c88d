#!/bin/bash
# Round-1 (e) evidence run on one B200: GPU test suite, smoke, bench lines, ncu launch lists of the bench command and
# --set full captures of trace_kernel_pb (reduced configs so the ~40 replays finish).  Every ncu pass follows a plain
# run of the same command that exited 0.
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_r01e.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_r01e.log; tail -3 $O/pytest_r01e.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r01e.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_r01e.log
python bench.py > $O/bench_r01e_cfg4.json 2> $O/bench_r01e_cfg4.err; echo "bench cfg4 rc=$?"
python bench.py --workload cfg2 --steps 10 --warmup 3 > $O/bench_r01e_cfg2.json 2> $O/bench_r01e_cfg2.err; echo "bench cfg2 rc=$?"
python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/bench_r01e_cfg5.json 2> $O/bench_r01e_cfg5.err; echo "bench cfg5 rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_r01e_ref.json 2> $O/bench_r01e_ref.err; echo "bench ref rc=$?"
FAST="--no-cpu-baseline --no-ref-gpu --no-lbvh-extra"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
python bench.py --workload cfg2 --steps 3 --warmup 1 $FAST > /dev/null 2>&1 && \
ncu --metrics $M --clock-control none --csv --log-file $O/r01e_bench_cfg2_launches.csv python bench.py --workload cfg2 --steps 3 --warmup 1 $FAST > $O/ncu_e_c2.log 2>&1
python bench.py --steps 1 --warmup 1 $FAST > /dev/null 2>&1 && \
ncu --metrics $M --clock-control none --csv --log-file $O/r01e_bench_cfg4_launches.csv python bench.py --steps 1 --warmup 1 $FAST > $O/ncu_e_c4.log 2>&1
B=raytracingincuda_b200/bin/b200-raytrace
CLI="$B --scene_id 1 --width 1920 --height 1080 --samples 16 --bounces 25 --no-ppm --stats"
$CLI > $O/plain_pb.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:trace_kernel_pb -o $O/prof_r01e_pb -f $CLI > $O/ncu_pb.log 2>&1
echo "ncu linear rc=$?"
$CLI --accel lbvh > $O/plain_pb_l1.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:trace_kernel_pb -o $O/prof_r01e_pb_lbvh_s1 -f $CLI --accel lbvh > $O/ncu_pb_l1.log 2>&1
echo "ncu lbvh s1 rc=$?"
$CLI --scaled_half 158 > $O/plain_pb_l2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:trace_kernel_pb -o $O/prof_r01e_pb_lbvh_100k -f $CLI --scaled_half 158 > $O/ncu_pb_l2.log 2>&1
echo "ncu lbvh 100k rc=$?"
ls -la $O/prof_r01e_pb*.ncu-rep
