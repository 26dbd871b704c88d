"""Jobs per pixel at config 4 (and config 5), default path: is one sample per job better there too?  usage: python tools/tune_plan3.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
import torch
r = rt.Renderer(0)
out = torch.empty((2160, 3840, 3), dtype=torch.float32, device="cuda:0")
res = {}
for name, slots, spp in (("cfg4", rt.scene(1), 1000), ("cfg5", rt.scene_scaled(158), 256)):
    r.upload_scene(slots)
    cam = rt.camera(3840, 2160, spp, 50)
    ms = []
    for k in range(3):
        r.render(cam, out=out)
        ms.append(r.stats().trace_ms)
    res[name] = (round(min(ms[1:]), 1), r.stats().chunks)
print(json.dumps(res))
'''
for chunks in ("8", "33", "125", "250", "500", "1000"):
    env = dict(os.environ, RT_CHUNKS=chunks)
    p = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True)
    print("chunks", chunks, p.stdout.strip() or p.stderr[-200:], flush=True)
