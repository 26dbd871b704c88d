for C in 1 4 8 12 16 24; do echo "RT_PB_COHORT=$C"; RT_PB_COHORT=$C python tools/tune.py --accel grid xx 2>/dev/null; RT_PB_COHORT=$C python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
import torch
r = rt.Renderer(0)
out = torch.empty((1080, 1920, 3), dtype=torch.float32, device="cuda:0")
res = {}
for name, slots, spp, depth, accel in (("scene1 grid", rt.scene(1), 100, 25, api.ACCEL_GRID), ("scene1 linear", rt.scene(1), 100, 25, api.ACCEL_LINEAR),
                                       ("100k lbvh", rt.scene_scaled(158), 32, 50, api.ACCEL_LBVH)):
    r.upload_scene(slots)
    cam = rt.camera(1920, 1080, spp, depth)
    ms = []
    for _ in range(4):
        r.render(cam, api.make_opts(accel=accel), out=out)
        ms.append(r.stats().trace_ms)
    res[name] = round(min(ms[1:]), 2)
print(res, flush=True)
PY
done
