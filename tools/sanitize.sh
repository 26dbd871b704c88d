#!/bin/bash
# compute-sanitizer pass over the render path: memcheck, racecheck, initcheck and synccheck on small
# renders of every kernel family (linear scan float/double, LBVH, uniform grid, wavefront, multi-chunk, primary pass).
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh'      -> gpurun_out/sanitize_<tool>.log, summary in gpurun_out/sanitize_summary.txt
set -u
export RT_ENABLE_GRID=1
O=${1:-gpurun_out}
mkdir -p "$O"
B=raytracingincuda_b200/bin/b200-raytrace
CS=/usr/local/cuda/bin/compute-sanitizer
: > "$O/sanitize_summary.txt"
run() {   # tool, label, command...
  local tool=$1 label=$2; shift 2
  local log="$O/sanitize_${tool}.log"
  echo "### $label: $*" >> "$log"
  timeout 600 $CS --tool "$tool" --error-exitcode 99 --print-limit 20 "$@" >> "$log" 2>&1
  local rc=$?
  local line
  line=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
  echo "$tool | $label | rc=$rc | ${line:-no summary}" | tee -a "$O/sanitize_summary.txt"
}
for tool in memcheck racecheck initcheck synccheck; do
  : > "$O/sanitize_${tool}.log"
  run $tool "linear float (scan + bins), scene 1"      $B --scene_id 1 --width 96 --height 64 --samples 9 --bounces 25 --no-ppm --accel linear
  run $tool "linear float, bins off"                   $B --scene_id 3 --width 64 --height 40 --samples 9 --bounces 50 --no-ppm --accel linear --primary_bins off
  run $tool "linear double, scene 2"                   $B --scene_id 2 --width 64 --height 40 --samples 9 --bounces 50 --no-ppm --accel linear --precision double
  run $tool "lbvh float, scene 1"                      $B --scene_id 1 --width 96 --height 64 --samples 9 --bounces 25 --no-ppm --accel lbvh
  run $tool "lbvh float, scaled scene (2 308 slots)"   $B --scene_id 1 --scaled_half 24 --width 96 --height 64 --samples 4 --bounces 25 --no-ppm
  run $tool "grid float, scene 1"                      $B --scene_id 1 --width 96 --height 64 --samples 9 --bounces 25 --no-ppm --accel grid
  run $tool "wavefront float, scene 1"                 $B --scene_id 1 --width 64 --height 40 --samples 9 --bounces 25 --no-ppm --kernel wavefront --accel linear
done
# the Python entry point: smoke() (render + primary pass through ctypes)
for tool in memcheck initcheck; do
  run $tool "smoke()" python -c "import __graft_entry__ as g; g.smoke()"
done
cat "$O/sanitize_summary.txt"
