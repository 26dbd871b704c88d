#!/bin/bash
# Final evidence run of round 2 on one B200 (supersedes tools/profile_r02.sh: same passes, but the captures are summarised into
# profiles/ ON THE BOX before the bench line is taken, so the line's roofline.traffic and roofline.ncu blocks describe the library
# build that produced it).  Every ncu pass follows a plain run of the same command that exited 0.  usage: bash tools/profile_r02_final.sh [tag]
set -u
O=gpurun_out
T=${1:-r02f}
python -m pytest tests -m gpu -x -q > $O/pytest_$T.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$T.log; tail -3 $O/pytest_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$T.log
python tools/time_accels.py 2>&1 | tee $O/${T}_time_accels.log
FAST="--no-cpu-baseline --no-ref-gpu --no-extras"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
LIBID=$(python -c "import bench; print(bench.lib_id())")
for WL in cfg2 cfg4; do
  STEPS=3; [ $WL = cfg4 ] && STEPS=1
  for ACC in auto linear; do
    python bench.py --workload $WL --accel $ACC --steps $STEPS --warmup 1 $FAST > /dev/null 2>&1 && \
    ncu --metrics $M --clock-control none --csv --log-file $O/${T}_bench_${WL}_${ACC}_launches.csv python bench.py --workload $WL --accel $ACC --steps $STEPS --warmup 1 $FAST > $O/ncu_${T}_${WL}_${ACC}.log 2>&1
    echo "launch list $WL $ACC rc=$?"
  done
done
python tools/ncu_launch_summary.py --lib-id=$LIBID cfg2=$O/${T}_bench_cfg2_auto_launches.csv cfg4=$O/${T}_bench_cfg4_auto_launches.csv \
    cfg2_linear=$O/${T}_bench_cfg2_linear_launches.csv cfg4_linear=$O/${T}_bench_cfg4_linear_launches.csv > $O/${T}_bench_kernel_traffic.json
B=raytracingincuda_b200/bin/b200-raytrace
CLI="$B --scene_id 1 --width 1920 --height 1080 --samples 16 --bounces 25 --no-ppm --stats"
for V in "linear:--accel linear:pbIfLi0" "grid:--accel grid:pbIfLi4" "lbvh_s1:--accel lbvh:pbIfLi3" "lbvh_100k:--scaled_half 158 --accel lbvh:pbIfLi1" "grid_100k:--scaled_half 158 --accel grid:pbIfLi4"; do
  NAME=${V%%:*}; REST=${V#*:}; FLAGS=${REST%%:*}; KN=${REST#*:}
  $CLI $FLAGS > $O/plain_${T}_$NAME.log 2>&1 && \
  ncu --set full --import-source on --clock-control none -k regex:trace_kernel_pb -o $O/prof_${T}_pb_$NAME -f $CLI $FLAGS > $O/ncu_${T}_$NAME.log 2>&1
  echo "ncu $NAME rc=$?"
  bash tools/ncu_summarise.sh $O/prof_${T}_pb_$NAME.ncu-rep $O/r02_pb_$NAME $KN && cp $O/r02_pb_${NAME}_* profiles/
done
echo "{\"lib_id\": \"$LIBID\"}" > $O/r02_capture_id.json
cp $O/r02_capture_id.json profiles/r02_capture_id.json
cp $O/${T}_bench_kernel_traffic.json profiles/r02_bench_kernel_traffic.json
for WL in cfg2 cfg4; do for ACC in auto linear; do cp $O/${T}_bench_${WL}_${ACC}_launches.csv $O/r02_bench_${WL}_${ACC}_launches.csv; done; done
ls -la $O/prof_${T}_pb*.ncu-rep
mkdir -p $O/benchmarks
python tools/benchmark.py --exe $B --out $O/benchmarks/b200_float.csv --scenes 1 --sizes 320x192,480x288,640x384,960x576,1280x768,1920x1080 --samples 10,100 --bounces 25 --threads 8 --runs 3 > /dev/null 2>&1
[ "${REF_SWEEP:-1}" = 1 ] && python tools/benchmark.py --exe oracle/_ref/global-float-cuda-raytrace --out $O/benchmarks/reference_global_float_sm100.csv --scenes 1 --sizes 320x192,480x288,640x384,960x576,1280x768,1920x1080 --samples 10,100 --bounces 25 --threads 8 --runs 3 > /dev/null 2>&1
echo "benchmark sweeps rc=$?"; cat $O/benchmarks/b200_float_avg.csv | head -20
python bench.py > $O/bench_${T}_cfg4.json 2> $O/bench_${T}_cfg4.err; echo "bench cfg4 rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_${T}_ref.json 2> $O/bench_${T}_ref.err; echo "bench ref rc=$?"
