"""Registers / spills of the trace kernels from csrc/ptxas.log.  usage: python tools/regs.py [substring]"""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = open(os.path.join(ROOT, "raytracingincuda_b200", "csrc", "ptxas.log")).read()
ents = re.findall(r"Compiling entry function '([^']+)'.*?\n(?:.*\n)*?ptxas info\s+: Used (\d+) registers[^\n]*", log)
names = subprocess.run(["c++filt"] + [e[0] for e in ents], capture_output=True, text=True).stdout.split("\n")
want = sys.argv[1] if len(sys.argv) > 1 else "trace_kernel_pb"
for (m, r), n in zip(ents, names):
    if want in n:
        blk = log[log.index(m):]
        spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", blk)
        print(r, spill.group(0) if spill else "", n[:70])
