# A/B of library variants under build/variants in ONE GPU call (same box, same clocks); AB_GLOB selects the libraries
python - <<'PY'
import glob, json, os, subprocess, sys
ROOT = os.getcwd()
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
import torch
r = rt.Renderer(0)
out = torch.empty((1080, 1920, 3), dtype=torch.float32, device="cuda:0")
res = {}
for name, slots, spp, depth, accel in (("scene1 grid", rt.scene(1), 100, 25, api.ACCEL_GRID), ("scene1 linear", rt.scene(1), 100, 25, api.ACCEL_LINEAR),
                                       ("scene1 lbvh", rt.scene(1), 100, 25, api.ACCEL_LBVH), ("scene3 grid", rt.scene(3), 100, 50, api.ACCEL_GRID), ("100k lbvh", rt.scene_scaled(158), 32, 50, api.ACCEL_LBVH),
                                       ("100k grid", rt.scene_scaled(158), 32, 50, api.ACCEL_GRID), ("14k grid", rt.scene_scaled(60), 64, 50, api.ACCEL_GRID)):
    r.upload_scene(slots)
    cam = rt.camera(1920, 1080, spp, depth)
    ms = []
    for _ in range(4):
        r.render(cam, api.make_opts(accel=accel), out=out)
        ms.append(r.stats().trace_ms)
    st = r.stats()
    res[name] = (round(min(ms[1:]), 2), round(st.binned_segments / st.paths, 3))
print(json.dumps(res))
'''
for lib in sorted(glob.glob(os.environ.get("AB_GLOB", "build/variants/librt_b200_*.so"))):
    env = dict(os.environ, RT_B200_LIB=os.path.abspath(lib))
    p = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), p.stdout.strip() or p.stderr[-300:], flush=True)
PY
