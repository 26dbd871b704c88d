# Grid on large fields: ring steps test only their leading edge (RT_GRID_RING_EDGE, rt_grid.cuh) against the whole block every step;
# every grid frame is compared with the LBVH's, bit for bit.
# usage: tools/build_variants.sh "ring1:-DRT_GRID_RING_EDGE=1" "ring0:-DRT_GRID_RING_EDGE=0"; bash tools/ab_ring.sh
python - <<'PY'
import json, os, subprocess, sys
ROOT = os.getcwd()
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
import torch
r = rt.Renderer(0)
res = {}
for name, slots, W, H, spp, depth in (("scene1 1080p", rt.scene(1), 1920, 1080, 100, 25), ("14k 1080p", rt.scene_scaled(60), 1920, 1080, 64, 50),
                                      ("100k 1080p", rt.scene_scaled(158), 1920, 1080, 32, 50), ("100k 4K", rt.scene_scaled(158), 3840, 2160, 32, 50),
                                      ("360k 1080p", rt.scene_scaled(300), 1920, 1080, 16, 50)):
    r.upload_scene(slots)
    cam = rt.camera(W, H, spp, depth)
    out = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
    ref = torch.empty_like(out)
    r.render(cam, api.make_opts(accel=api.ACCEL_LBVH), out=ref)
    lb = r.stats().trace_ms
    ms = []
    for _ in range(3):
        r.render(cam, api.make_opts(accel=api.ACCEL_GRID), out=out)
        ms.append(r.stats().trace_ms)
    st = r.stats()
    res[name] = (round(min(ms), 2), round(st.node_visits / st.segments, 2), round(st.sphere_tests / st.segments, 2), "lbvh %%.1f" %% lb,
                 "same" if torch.equal(out.view(torch.int32), ref.view(torch.int32)) else "DIFFERENT")
print(json.dumps(res))
'''
for lib in ("ring0", "ring1", "ring0", "ring1"):
    env = dict(os.environ, RT_B200_LIB=os.path.abspath(f"build/variants/librt_b200_{lib}.so"))
    p = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True)
    print(lib, p.stdout.strip() or p.stderr[-400:], flush=True)
PY
