#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS export with `nvdisasm -g` line info and report the
warp-instructions executed per source line (and per source function-ish region).

usage: ncu_by_line.py <source.csv> <nvdisasm -g output> <kernel mangled-name substring>
"""
import csv
import re
import sys
from collections import Counter

csv_path, dis_path, kname = sys.argv[1:4]
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
counts = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    counts.append((int(r[ix["Address"]], 16), int(r[ix["Instructions Executed"]]), r[ix["Source"]].strip()))
base = counts[0][0]

# walk the disassembly of the kernel: remember the current line tag, map instruction offset -> line
line_of = {}
cur = None
inside = False
for ln in open(dis_path):
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
by_line = Counter()
tot = 0
for addr, n, src in counts:
    by_line[line_of.get(addr - base)] += n
    tot += n
print("total", tot)
for (key, n) in by_line.most_common(60):
    print(f"{100*n/tot:6.2f}%  {key}")
