#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS export with `nvdisasm -gi` line info and report the share of
executed warp-instructions and of stall samples per source position.

usage: ncu_by_line.py <source.csv> <nvdisasm -gi output> <kernel mangled-name substring> [frames]

`frames` (default 1) = how many frames of the inlining chain, counted from the kernel body inwards, make up a
key: 1 attributes everything to the line of the kernel body that (transitively) called it, 2 adds the callee's
line, ... ; 0 uses the innermost line only.
"""
import csv
import re
import sys
from collections import Counter

csv_path, dis_path, kname = sys.argv[1:4]
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
counts = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    counts.append((int(r[ix["Address"]], 16), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]] or 0),
                   float(r[ix["Avg. Threads Executed"]] or 0)))
base = counts[0][0]

line_of = {}
chain = []
fresh = True
inside = False
for ln in open(dis_path):
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m:
        key = tuple(chain[-frames:][::-1]) if frames > 0 else tuple(chain[:1])
        line_of[int(m.group(1), 16)] = key
        fresh = True
by = Counter()
smp = Counter()
thr = Counter()
tot = ts = 0
for addr, n, s, t in counts:
    k = line_of.get(addr - base)
    by[k] += n
    smp[k] += s
    thr[k] += n * t
    tot += n
    ts += s
print("total warp-instructions", tot, "samples", ts)
for key, n in by.most_common(70):
    name = " <- ".join(f"{f}:{l}" for f, l in key) if key else "?"
    print(f"{100*n/tot:6.2f}% instr  {100*smp[key]/max(ts,1):6.2f}% samples  thr {thr[key]/max(n,1):5.1f}  {name}")
