"""Diagnostic build (-DRT_DIAG_LONG_TRAVERSAL=N, tools/build_variants.sh): print the rays whose LBVH traversal passes N, 2N, ... node
visits on config 5's scene (99 860 slots, 3840x2160, 32 spp).  usage: RT_B200_LIB=build/variants/librt_b200_diag.so python tools/diag_long_traversals.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
r = rt.Renderer(0)
r.upload_scene(rt.scene_scaled(158))
out = torch.empty((2160, 3840, 3), dtype=torch.float32, device="cuda:0")
cam = rt.camera(3840, 2160, 32, 50)
for k in range(2):
    r.render(cam, api.make_opts(accel=api.ACCEL_LBVH), out=out)
    torch.cuda.synchronize()
    st = r.stats()
    print(f"render {k}: trace_ms {st.trace_ms:.1f} nodes/seg {st.node_visits / st.segments:.2f}", flush=True)
