"""Times RT_ACCEL_GRID (experimental) next to the LBVH: scene 1 at config 2 and the 99 860-slot scene at 1080p / 32 spp.
usage: RT_ENABLE_GRID=1 python tools/time_grid.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
r = rt.Renderer(0)
for name, slots, spp, depth in (("scene1", rt.scene(1), 100, 25), ("100k", rt.scene_scaled(158), 32, 50)):
    r.upload_scene(slots)
    cam = rt.camera(1920, 1080, spp, depth)
    out = {}
    for label, accel in (("grid", api.ACCEL_GRID), ("lbvh", api.ACCEL_LBVH)):
        ms = []
        for _ in range(3):
            r.render(cam, api.make_opts(accel=accel))
            ms.append(r.stats().trace_ms)
        st = r.stats()
        out[label] = (round(min(ms), 2), round(st.node_visits / st.segments, 2), round(st.sphere_tests / st.segments, 2), st.regs)
    print(name, out, flush=True)
