"""Jobs per pixel x tail multiplier at config 2 / config 3 (short launches), default path.  usage: python tools/tune_plan2.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
import torch
r = rt.Renderer(0)
out = torch.empty((1080, 1920, 3), dtype=torch.float32, device="cuda:0")
res = {}
for name, sid, depth in (("cfg2", 1, 25), ("cfg3a", 2, 50)):
    r.upload_scene(rt.scene(sid))
    cam = rt.camera(1920, 1080, 100, depth)
    ms = []
    for k in range(4):
        r.render(cam, out=out)
        ms.append(r.stats().trace_ms)
    res[name] = round(min(ms[1:]), 2)
print(json.dumps(res))
'''
for chunks in ("4", "8", "12", "25", "50", "100"):
    for tail in ("1", "4", "8"):
        env = dict(os.environ, RT_CHUNKS=chunks, RT_TAIL_MULT=tail)
        p = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True)
        print("chunks", chunks, "tail x", tail, p.stdout.strip() or p.stderr[-200:], flush=True)
