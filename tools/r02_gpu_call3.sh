#!/bin/bash
# Round 2, third GPU call: the new bench line, the bench-schema test, scheduling knobs of the job plan, one ncu launch list.
set -u
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bench_line_schema or spot_check or auto_picks or grid_random or double" 2>&1 | tail -8 | tee $O/r02c_tests.log
python bench.py --steps 3 --warmup 3 > $O/r02c_bench_cfg4.json 2> $O/r02c_bench_cfg4.err
echo "bench rc=$?"
python tools/tune_plan.py 2>&1 | tee $O/r02c_tune_plan.log
