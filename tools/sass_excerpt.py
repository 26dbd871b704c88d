#!/usr/bin/env python
"""SASS evidence for the claims DESIGN.md makes about the shipped kernels (profiles/<tag>_sass_*.txt):
  * the scene blob reaches shared memory by a TMA bulk copy (UBLKCP) completed on an mbarrier (SYNCS ... TRYWAIT),
  * the scan body is 7 packed FFMA2 + 1 LDS.128 + 2 FSETP + 2 predicated IADD/LOP per record of two tests,
  * radiance is accumulated with 64-bit integer reductions (RED.E.ADD.64),
  * per-kernel opcode histogram.
usage: python tools/sass_excerpt.py [library.so] [out_prefix]"""
import collections
import hashlib
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "raytracingincuda_b200", "librt_b200.so")
prefix = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_sass")
KERNELS = {"linear": "_ZN2rt15trace_kernel_pbIfLi0EEEvNS_9TraceArgsIT_EE", "grid": "_ZN2rt15trace_kernel_pbIfLi4EEEvNS_9TraceArgsIT_EE",
           "lbvh": "_ZN2rt15trace_kernel_pbIfLi1EEEvNS_9TraceArgsIT_EE", "finalize": "_ZN2rt20finalize_flat_kernelIfEEvNS_10AccSourcesEyT_PS2_"}
import ctypes
_L = ctypes.CDLL(lib)
_L.rt_kernel_build_id.restype = ctypes.c_char_p
lib_id = _L.rt_kernel_build_id().decode()
for tag, fun in KERNELS.items():
    text = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
    ins = [l for l in text.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    ops = collections.Counter()
    for l in ins:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if m:
            full = m.group(1)
            ops[full if full.startswith(("RED", "UBLKCP", "SYNCS", "LDS", "ATOM")) else full.split(".")[0]] += 1
    with open(f"{prefix}_{tag}.txt", "w") as f:
        f.write(f"# cuobjdump -sass -fun {fun}\n# library {os.path.basename(lib)} lib_id {lib_id} (bench.py config.lib_id), arch sm_100a\n")
        f.write(f"# {len(ins)} instructions; opcode histogram (static):\n")
        for op, n in ops.most_common(28):
            f.write(f"#   {op:24s} {n}\n")
        def dump(title, pred, ctx=0, limit=40):
            f.write(f"\n## {title}\n")
            shown = 0
            for k, l in enumerate(ins):
                if pred(l) and shown < limit:
                    for q in ins[max(0, k - ctx):k + ctx + 1]:
                        f.write(q.rstrip() + "\n")
                    if ctx:
                        f.write("    ...\n")
                    shown += 1
        if tag != "finalize":
            dump("TMA bulk copy of the scene blob + mbarrier wait", lambda l: "UBLKCP" in l or "SYNCS" in l, 0, 12)
            dump("integer radiance accumulation (atomicAdd on int64 -> RED / ATOMG .ADD.64)", lambda l: ("RED" in l or "ATOMG" in l) and ".64" in l, 0, 12)
        if tag == "linear":
            # the first full unrolled filter block: from the first LDS.128 that is followed by FFMA2 to the mask store
            start = next(k for k, l in enumerate(ins) if "LDS.128" in l and any("FFMA2" in x for x in ins[k:k + 12]))
            f.write("\n## scan body: first 3 records of the unrolled block of 16 (LDS.128, 7 FFMA2, 2 FSETP, 2 predicated mask ORs each)\n")
            n_lds = 0
            for l in ins[start:start + 400]:
                if "LDS.128" in l:
                    n_lds += 1
                f.write(l.rstrip() + "\n")
                if sum(1 for x in ins[start:start + 400][:ins[start:start + 400].index(l) + 1] if "FFMA2" in x) >= 21:
                    break
            body = ins[start:start + 16 * 14]
            c = collections.Counter(re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l).group(1) for l in body)
            f.write(f"\n# opcode counts over the {len(body)} instructions that follow (about 16 records): {dict(c)}\n")
    print(f"{prefix}_{tag}.txt: {len(ins)} instructions")
