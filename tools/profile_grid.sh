#!/bin/bash
# one ncu --set full capture of the default-path kernel (grid) at the reduced config, after a plain run.  usage: bash tools/profile_grid.sh tag
set -u
O=gpurun_out; T=${1:-r02}
B=raytracingincuda_b200/bin/b200-raytrace
CLI="$B --scene_id 1 --width 1920 --height 1080 --samples 16 --bounces 25 --no-ppm --stats --accel grid"
$CLI > $O/plain_${T}_grid.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:trace_kernel_pb -o $O/prof_${T}_pb_grid -f $CLI > $O/ncu_${T}_grid.log 2>&1
echo "ncu grid rc=$?"; cat $O/plain_${T}_grid.log
cp raytracingincuda_b200/librt_b200.so $O/librt_b200_$T.so
