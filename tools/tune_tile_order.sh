for V in 0 1; do
  if [ $V = 1 ]; then export RT_NO_TILE_ORDER=1; echo "row-major numbering"; else unset RT_NO_TILE_ORDER; echo "8x4 tile numbering"; fi
  python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
import torch
r = rt.Renderer(0)
res = {}
for name, slots, W, H, spp, depth, accel in (("cfg2 grid", rt.scene(1), 1920, 1080, 100, 25, api.ACCEL_GRID), ("cfg2 linear", rt.scene(1), 1920, 1080, 100, 25, api.ACCEL_LINEAR),
                                             ("cfg3a grid", rt.scene(2), 1920, 1080, 100, 50, api.ACCEL_GRID),
                                             ("100k lbvh", rt.scene_scaled(158), 1920, 1080, 32, 50, api.ACCEL_LBVH), ("cfg4 grid", rt.scene(1), 3840, 2160, 1000, 50, api.ACCEL_GRID)):
    out = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
    r.upload_scene(slots)
    cam = rt.camera(W, H, spp, depth)
    ms = []
    for _ in range(3):
        r.render(cam, api.make_opts(accel=accel), out=out)
        ms.append(r.stats().trace_ms)
    res[name] = round(min(ms[1:]), 2)
print(res, flush=True)
PY
done
