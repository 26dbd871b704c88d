#!/bin/bash
# LBVH round shape: sweep RT_BVH_STEPS (max node visits per turn) and RT_BVH_MIN_ACTIVE (end the round
# once fewer lanes than this are still traversing) on a workload (run on a B200)
for w in cfg2 cfg5; do for s in 16 24 48; do for m in 0 12 16 20 24 28; do
  RT_BVH_STEPS=$s RT_BVH_MIN_ACTIVE=$m python bench.py --workload $w --accel lbvh --steps 2 --warmup 1 --no-cpu-baseline --no-ref-gpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w steps=$s min_active=$m', d['value'], d['ms_per_step'])"
done; done; done
