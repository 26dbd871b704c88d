#!/bin/bash
# Camera-ray bins with the LBVH: off vs on at BASELINE config 2 (scene 1 through the tree) and on the 99 860-slot
# scene of config 5 at a reduced sample count (run on a B200).
CLI=raytracingincuda_b200/bin/b200-raytrace
s1() { "$CLI" --scene_id 1 --width 1920 --height 1080 --samples 100 --bounces 25 --accel lbvh --no-ppm "$@" | tr -d ' ' | cut -d, -f1; }
s5() { "$CLI" --scene_id 1 --scaled_half 158 --width 3840 --height 2160 --samples 16 --bounces 50 --no-ppm "$@" | tr -d ' ' | cut -d, -f1; }
echo "scene 1 lbvh, bins off: $(s1 --primary_bins off) $(s1 --primary_bins off) ms"
echo "scene 1 lbvh, bins on : $(s1) $(s1) ms"
echo "100k lbvh, bins off: $(s5 --primary_bins off) $(s5 --primary_bins off) ms"
echo "100k lbvh, bins on : $(s5) $(s5) ms"
for st in 12 16 24; do for m in 0 6 12; do
  echo "scene1 steps=$st min_active=$m: $(RT_BVH_STEPS=$st RT_BVH_MIN_ACTIVE=$m s1) ms;  100k: $(RT_BVH_STEPS=$st RT_BVH_MIN_ACTIVE=$m s5) ms"
done; done
