#!/usr/bin/env python
"""Per-kernel summary of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch lists.  usage: ncu_launch_summary.py [--lib-id=ID] name=launches.csv [name=launches.csv ...] > traffic.json"""
import csv
import json
import sys
from collections import defaultdict

out = {}
args = sys.argv[1:]
if args and args[0].startswith("--lib-id="):          # hash of the library the launch lists were taken with (bench.py: lib_id)
    out["lib_id"] = args.pop(0).split("=", 1)[1]
for arg in args:
    name, path = arg.split("=", 1)
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0] != "ID"]
    per = defaultdict(lambda: defaultdict(dict))
    for r in rows:
        per[r[4]][r[0]][r[12]] = float(r[14]) * (1e-6 if r[13] == "ns" else 1.0)
    tot = sum(m.get("gpu__time_duration.sum", 0.0) for k in per.values() for m in k.values())
    summ = {}
    for kname, launches in per.items():
        short = kname.split("(")[0].replace("rt::", "")
        n = len(launches)
        ms = sum(m.get("gpu__time_duration.sum", 0.0) for m in launches.values())
        summ[short] = {
            "launches": n, "mean_ms": ms / n,
            "dram_read_bytes_per_launch": sum(m.get("dram__bytes_read.sum", 0.0) for m in launches.values()) / n,
            "dram_write_bytes_per_launch": sum(m.get("dram__bytes_write.sum", 0.0) for m in launches.values()) / n,
            "share_of_gpu_time": round(ms / tot, 5) if tot else None,
        }
    out[name] = summ
json.dump(out, sys.stdout, indent=1)
print()
