#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 / "transient").  usage: tools/gpu_retry.sh <log> [gpurun args...] -- 'command'
LOG=$1; shift
for try in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$LOG"; then echo "gpu_retry: done rc=$rc try=$try" >> "$LOG"; exit $rc; fi
  sleep 90
done
echo "gpu_retry: gave up" >> "$LOG"
