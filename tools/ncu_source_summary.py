#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: instruction mix by opcode class and the
hottest instructions.  Usage: ncu -i rep --page source --csv | python tools/ncu_source_summary.py"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(sys.stdin))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot = 0
by_op = Counter()
thr = Counter()
insts = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    n = int(r[ix["Instructions Executed"]])
    t = int(r[ix["Thread Instructions Executed"]])
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    by_op[op] += n
    thr[op] += t
    tot += n
    insts.append((n, t, src))
print("total warp-instructions", tot)
FMA = {"FFMA", "FFMA2", "FMUL", "FMUL2", "FADD", "FADD2", "IMAD", "HFMA2", "FFMA32I", "FMUL32I", "FADD32I"}
print("fma-pipe share %.3f" % (sum(v for k, v in by_op.items() if k in FMA) / tot))
for op, n in by_op.most_common(40):
    print(f"{op:10s} {n:14d} {100*n/tot:6.2f}%  avg-threads {thr[op]/max(1,n):5.1f}")
