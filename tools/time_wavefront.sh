#!/bin/bash
# wavefront variant (paired filter scan since round 2) against the megakernel on the same scan, render_ms of the CLI, best of 3
B=raytracingincuda_b200/bin/b200-raytrace
for S in 1 2 3; do
  for K in "mega --accel linear --primary_bins off" "mega --accel linear" "wavefront --accel linear" "mega --accel auto"; do
    best=999999
    for r in 1 2 3; do
      ms=$($B --scene_id $S --width 1920 --height 1080 --samples 100 --bounces 50 --no-ppm --kernel $K | cut -d, -f1 | tr -d ' ')
      best=$(python -c "print(min($best, $ms))")
    done
    echo "scene $S --kernel $K: $best ms"
  done
done
