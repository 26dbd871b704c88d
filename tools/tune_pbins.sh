#!/bin/bash
# Camera-ray bins (rt_primary_bins.cuh): off vs on, and the round shape of phase A (RT_PB_ROUNDS camera-ray rounds per
# loop turn, later rounds need RT_PB_MIN fresh lanes), on BASELINE config 2 (run on a B200).
CLI=raytracingincuda_b200/bin/b200-raytrace
run() { "$CLI" --scene_id 1 --width 1920 --height 1080 --samples 100 --bounces 25 --no-ppm "$@" | tr -d ' ' | cut -d, -f1; }
echo "bins off: $(run --primary_bins off) $(run --primary_bins off) ms"
for r in 1 2 3 4 6; do for m in 1 4 8 12; do
  echo "rounds=$r min=$m: $(RT_PB_ROUNDS=$r RT_PB_MIN=$m run) $(RT_PB_ROUNDS=$r RT_PB_MIN=$m run) ms"
done; done
