#!/usr/bin/env python
"""Static SASS size of one kernel by source position (nvdisasm -gi line info): where the code footprint is -- the traversal
kernels are instruction-fetch bound (profiles/README.md), so inlined copies matter.
usage: python tools/sass_footprint.py <mangled-name substring> [library.so]"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kn = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "raytracingincuda_b200", "librt_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", cubin], cwd=tmp, capture_output=True, text=True).stdout.split("\n")
inside, chain, fresh, total = False, [], True, 0
inner, outer = collections.Counter(), collections.Counter()
for ln in dis:
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        inside = kn in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln):
        total += 1
        inner[chain[0] if chain else ("?", 0)] += 1
        outer[chain[-1] if chain else ("?", 0)] += 1
        fresh = True
print("total instructions", total)
print("by kernel-body line:")
for k, v in outer.most_common(14):
    print(f"  {v:5d} {k[0]}:{k[1]}")
print("by innermost line:")
for k, v in inner.most_common(22):
    print(f"  {v:5d} {k[0]}:{k[1]}")
