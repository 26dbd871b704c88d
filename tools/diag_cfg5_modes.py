"""Config 5 (99 860 slots, 3840x2160): is the LBVH render time per sample stable?  The 8-GPU bench line shows ranks at 157 and at
249 ms for the same 32-spp share.  Repeated renders in one process, then fresh processes, LBVH and grid, 32 and 128 spp; device
time of the trace phase (tile lists + trace kernel) next to the wall time of the call.
usage: python tools/diag_cfg5_modes.py [child]"""
import os, subprocess, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def child():
    import torch
    import raytracingincuda_b200 as rt
    from raytracingincuda_b200 import api
    r = rt.Renderer(0)
    r.upload_scene(rt.scene_scaled(158))
    out = torch.empty((2160, 3840, 3), dtype=torch.float32, device="cuda:0")
    for label, accel in (("lbvh", api.ACCEL_LBVH), ("grid", api.ACCEL_GRID)):
        for spp, reps in ((32, 6), (128, 2)):
            cam = rt.camera(3840, 2160, spp, 50)
            res = []
            for _ in range(reps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r.render(cam, api.make_opts(accel=accel), out=out)
                torch.cuda.synchronize()
                wall = (time.perf_counter() - t0) * 1e3
                st = r.stats()
                res.append((round(st.trace_ms, 1), round(st.render_ms, 1), round(wall, 1)))
            print(f"  {label} {spp:3d} spp  (trace_ms, render_ms, wall_ms): {res}  nodes/seg {st.node_visits / st.segments:.2f}", flush=True)
        # the spp split's shares: does the sample range matter?
        cam = rt.camera(3840, 2160, 256, 50)
        acc = torch.empty((2160 * 3840 * 3,), dtype=torch.int64, device="cuda:0")
        res = []
        for rank in (0, 2, 5, 7):
            o = api.make_opts(accel=accel, split=api.SPLIT_SPP, rank=rank, world=8)
            r.render_partials(cam, o, acc)
            res.append((rank, round(r.stats().trace_ms, 1)))
        print(f"  {label} spp-split shares of 256 spp (rank, trace_ms): {res}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for k in range(3):
            print(f"process {k}", flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], check=False)
