#!/bin/bash
# LBVH node visits per loop turn: sweep RT_BVH_STEPS on a workload (run on a B200)
for w in cfg2 cfg5; do for s in 4 6 8 12 16 24 32 48; do
  RT_BVH_STEPS=$s python bench.py --workload $w --accel lbvh --steps 2 --warmup 1 --no-cpu-baseline --no-ref-gpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w steps=$s', d['value'], d['ms_per_step'])"
done; done
