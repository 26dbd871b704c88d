#!/bin/bash
# Turn one .ncu-rep (ncu --set full --import-source on) into the four text summaries kept under profiles/:
#   <out>_ncu_details.txt, <out>_key_metrics.csv, <out>_instruction_mix.txt, <out>_by_source_line.txt
# usage: tools/ncu_summarise.sh <report.ncu-rep> <profiles/prefix> <mangled-kernel-name substring> [inner-frame file:line filter]
# The by-line table needs the cubin the capture ran (same build of librt_b200.so).
set -e
REP=$1; OUT=$2; KN=$3
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
KEYS='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|l1tex__t_sector_hit_rate.pct|launch__block_size|launch__grid_size|launch__registers_per_thread|lts__t_sector_hit_rate.pct|sm__inst_executed.avg.per_cycle_active|sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active|sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum|smsp__issue_active.avg.pct_of_peak_sustained_active|smsp__sass_average_branch_targets_threads_uniform.pct|smsp__sass_branch_targets_threads_divergent.sum|smsp__thread_inst_executed_per_inst_executed.ratio'
ncu -i "$REP" --page details > "${OUT}_ncu_details.txt"
ncu -i "$REP" --page raw --csv > $TMP/raw.csv
python - "$TMP/raw.csv" "$KEYS" > "${OUT}_key_metrics.csv" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, val = rows[0], rows[1], rows[2]
keys = set(sys.argv[2].split("|"))
for h, u, v in sorted(zip(hdr, units, val)):
    if h in keys:
        print(f"{h},{u},{v}")
PY
ncu -i "$REP" --page source --csv > $TMP/src.csv
python $ROOT/tools/ncu_source_summary.py < $TMP/src.csv > "${OUT}_instruction_mix.txt"
(cd $TMP && cuobjdump -xelf all $ROOT/raytracingincuda_b200/librt_b200.so > /dev/null && nvdisasm -gi rt_kernels.sm_100a.cubin > dis.txt 2>/dev/null)
{
  echo "# share of executed warp-instructions / stall samples by line of the kernel body that the code was inlined from"
  python $ROOT/tools/ncu_by_line.py $TMP/src.csv $TMP/dis.txt "$KN" 1 2>/dev/null | head -24
  echo
  echo "# two frames of the inlining chain (kernel body line <- callee line)"
  python $ROOT/tools/ncu_by_line.py $TMP/src.csv $TMP/dis.txt "$KN" 2 2>/dev/null | head -32
} > "${OUT}_by_source_line.txt"
rm -rf $TMP
