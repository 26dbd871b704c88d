#!/usr/bin/env python
"""Time tuning variants of the library (build/variants/librt_b200_*.so) on one workload.
usage: python tools/tune.py [--workload cfg2] [--accel auto|linear|lbvh|grid] [name ...]"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import raytracingincuda_b200 as rt
scene, W, H, spp, depth = %r
from raytracingincuda_b200 import api
ACCEL = {"auto": api.ACCEL_AUTO, "linear": api.ACCEL_LINEAR, "lbvh": api.ACCEL_LBVH, "grid": api.ACCEL_GRID}[%r]
r = rt.Renderer(0)
r.upload_scene(rt.scene(scene))
cam = rt.camera(W, H, spp, depth)
import torch
out = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
ms = []
for k in range(5):
    r.render(cam, api.make_opts(accel=ACCEL), out=out)
    ms.append(r.stats().trace_ms)
st = r.stats()
print(json.dumps({"ms": min(ms[1:]), "regs": st.regs, "grid": st.grid, "mps": W*H*spp/min(ms[1:])/1e3}))
'''
WORK = {"cfg2": (1, 1920, 1080, 100, 25), "cfg3a": (2, 1920, 1080, 100, 50), "cfg3b": (3, 1920, 1080, 100, 50),
        "small": (1, 320, 192, 100, 25)}


def main():
    args = sys.argv[1:]
    wl, accel = "cfg2", "auto"
    while args and args[0] in ("--workload", "--accel"):
        if args[0] == "--workload":
            wl = args[1]
        else:
            accel = args[1]
        args = args[2:]
    libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "librt_b200_*.so")))
    if args:
        libs = [l for l in libs if any(a in os.path.basename(l) for a in args)]
    for lib in libs:
        env = dict(os.environ, RT_B200_LIB=lib)
        p = subprocess.run([sys.executable, "-c", CHILD % (ROOT, WORK[wl], accel)], env=env, capture_output=True, text=True)
        name = os.path.basename(lib)[len("librt_b200_"):-3]
        print(name, wl, accel, p.stdout.strip() or p.stderr.strip()[-300:], flush=True)


if __name__ == "__main__":
    main()
