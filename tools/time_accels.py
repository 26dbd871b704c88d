"""Every structure hit_world can use, timed on BASELINE's scenes at 1920x1080 / 100 spp: what should RT_ACCEL_AUTO pick?
usage: python tools/time_accels.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
r = rt.Renderer(0)
out = torch.empty((1080, 1920, 3), dtype=torch.float32, device="cuda:0")
for name, slots, spp, depth in (("scene2 (40)", rt.scene(2), 100, 50), ("scene3 (125)", rt.scene(3), 100, 50), ("scene1 (488)", rt.scene(1), 100, 25),
                                ("scaled12 (580)", rt.scene_scaled(12), 100, 50), ("scaled30 (3604)", rt.scene_scaled(30), 100, 50),
                                ("scaled60 (14404)", rt.scene_scaled(60), 64, 50), ("scaled158 (99860)", rt.scene_scaled(158), 32, 50)):
    r.upload_scene(slots)
    cam = rt.camera(1920, 1080, spp, depth)
    res = {}
    for label, accel in (("linear", api.ACCEL_LINEAR), ("lbvh", api.ACCEL_LBVH), ("grid", api.ACCEL_GRID), ("auto", api.ACCEL_AUTO)):
        if label == "linear" and len(slots) > 5000:
            continue
        ms = []
        try:
            for _ in range(3):
                r.render(cam, api.make_opts(accel=accel), out=out)
                ms.append(r.stats().trace_ms)
            st = r.stats()
            res[label] = (round(min(ms), 2), api.ACCEL_NAMES[st.accel_used], round(st.node_visits / st.segments, 2), round(st.sphere_tests / st.segments, 2))
        except rt.RtError as e:
            res[label] = str(e)[-40:]
    print(name, res, flush=True)
