#!/bin/bash
# BASELINE config 3: scenes 2 and 3, 1920x1080, 100 spp, 50 bounces, float and double --
# the new binary next to the reference binaries rebuilt for sm_100 (oracle/_ref).  Run on a B200.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "${1:-.}"
B=$ROOT/raytracingincuda_b200/bin/b200-raytrace
A="--width 1920 --height 1080 --samples 100 --bounces 50 --threads 8"
echo "impl,precision,scene,render_ms,e2e_ms"
for s in 1 2 3; do
  for p in float double; do
    echo "b200,$p,$s,$($B --scene_id $s $A --precision $p --no-ppm | tr -d ' ')"
    echo "b200,$p,$s,$($B --scene_id $s $A --precision $p --no-ppm | tr -d ' ')"
    echo "reference-sm100,$p,$s,$($ROOT/oracle/_ref/global-$p-cuda-raytrace --scene_id $s $A | tr -d ' ')"
  done
done
