#!/usr/bin/env python
"""Time the LBVH kernel of library variants (build/variants/) on scene 1 (compact-scene variant) and on the
99 860-slot scene.  usage: python tools/tune_lbvh.py"""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
import torch
r = rt.Renderer(0)
res = {}
for name, slots, spp in (("scene1", rt.scene(1), 100), ("100k", rt.scene_scaled(158), 32)):
    r.upload_scene(slots)
    W, H = 1920, 1080
    cam = rt.camera(W, H, spp, 50)
    out = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
    ms = []
    for k in range(4):
        r.render(cam, api.make_opts(accel=api.ACCEL_LBVH), out=out)
        ms.append(r.stats().trace_ms)
    st = r.stats()
    res[name] = {"ms": round(min(ms[1:]), 3), "mps": round(W*H*spp/min(ms[1:])/1e3, 1), "regs": st.regs,
                 "nodes_per_seg": round(st.node_visits/st.segments, 2), "tests_per_seg": round(st.sphere_tests/st.segments, 2)}
print(json.dumps(res))
'''
libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "librt_b200_*.so"))) or [os.path.join(ROOT, "raytracingincuda_b200", "librt_b200.so")]
for lib in libs:
    p = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=dict(os.environ, RT_B200_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib), p.stdout.strip() or p.stderr.strip()[-300:], flush=True)
