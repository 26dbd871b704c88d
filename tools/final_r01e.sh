#!/bin/bash
# Final single-GPU run of round 1 (e): GPU suite, smoke, bench lines (configs 4, 2, 5), config-3 table.
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_r01e.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_r01e.log; tail -3 $O/pytest_r01e.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r01e.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_r01e.log
python bench.py > $O/bench_r01e_cfg4.json 2> $O/bench_r01e_cfg4.err; echo "bench cfg4 rc=$?"
python bench.py --workload cfg2 --steps 10 --warmup 3 > $O/bench_r01e_cfg2.json 2> $O/bench_r01e_cfg2.err; echo "bench cfg2 rc=$?"
python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/bench_r01e_cfg5.json 2> $O/bench_r01e_cfg5.err; echo "bench cfg5 rc=$?"
bash tools/compare_cfg3.sh /tmp > $O/cfg3_r01e.csv 2>&1; cat $O/cfg3_r01e.csv
