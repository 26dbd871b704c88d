"""Scheduling knobs of the job plan (plan_jobs, rt_kernels.cu) at BASELINE config 4 and config 2, linear scan and default path:
band height (L2 locality of the accumulators), jobs per pixel and the finer tail region (drain at the end of the launch).
The image does not depend on any of them (integer accumulation); only the time does.
usage: python tools/tune_plan.py [cfg4|cfg2 ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
import torch
scene, W, H, spp, depth, accel = %r
r = rt.Renderer(0)
r.upload_scene(rt.scene(scene))
cam = rt.camera(W, H, spp, depth)
out = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
ms = []
for k in range(3):
    r.render(cam, api.make_opts(accel=accel), out=out)
    ms.append(r.stats().trace_ms)
st = r.stats()
print(json.dumps({"ms": round(min(ms[1:]), 2), "chunks": st.chunks, "mps": round(W*H*spp/min(ms[1:])/1e3, 1)}))
'''
WORK = {"cfg4": (1, 3840, 2160, 1000, 50), "cfg2": (1, 1920, 1080, 100, 25)}
KNOBS = [{}, {"RT_TAIL_MULT": "1"}, {"RT_TAIL_MULT": "8"}, {"RT_BAND_ROWS": "16"}, {"RT_BAND_ROWS": "256"}, {"RT_BAND_ROWS": "100000"},
         {"RT_CHUNKS": "8"}, {"RT_CHUNKS": "64"}, {"RT_CHUNKS": "125"}]


def main():
    names = sys.argv[1:] or ["cfg4", "cfg2"]
    for wl in names:
        for accel, label in ((0, "linear"), (2, "auto")):
            for knobs in KNOBS:
                if wl == "cfg4" and accel == 2 and knobs and "RT_TAIL_MULT" not in knobs:
                    continue
                env = dict(os.environ, **knobs)
                p = subprocess.run([sys.executable, "-c", CHILD % (ROOT, WORK[wl] + (accel,))], env=env, capture_output=True, text=True)
                print(wl, label, knobs or "default", p.stdout.strip() or p.stderr.strip()[-300:], flush=True)


if __name__ == "__main__":
    main()
