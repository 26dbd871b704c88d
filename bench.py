#!/usr/bin/env python
"""bench.py -- headline benchmark of the render hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full render of the workload (default: BASELINE config 4, scene 1, 3840x2160,
1000 spp, 50 bounces, float).  `value` = W*H*spp*1e-6 / step-seconds (Mpath-samples/s), timed with
CUDA events on the launching stream, scene already resident in HBM, max over ranks.  `e2e` is the
same metric through the public C-ABI call with HOST buffers: per step the scene slots are uploaded
from host memory and the gamma-encoded frame is read back into pinned host memory.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene, width, height, spp, bounces)
    "cfg4": (1, 3840, 2160, 1000, 50),      # BASELINE.json configs[3]: the config the metric is quoted on
    "cfg2": (1, 1920, 1080, 100, 25),       # configs[1]
    "cfg1": (1, 320, 192, 10, 25),          # configs[0] (the reference's own CPU-runnable case)
    "cfg3a": (2, 1920, 1080, 100, 50),
    "cfg3b": (3, 1920, 1080, 100, 50),
    # configs[4]: scaled random-spheres scene (grid [-158,158)^2 -> 99 860 slots), on-GPU LBVH; scene id = -half
    "cfg5": (-158, 3840, 2160, 256, 50),
}
METRIC = "Mpath-samples/s"
SLOTS = {1: 488, 2: 40, 3: 125}


def workload_name(name):
    """The same string in both arms' `config.workload`."""
    scene, W, H, spp, depth = WORKLOADS[name]
    what = (f"scaled random-spheres scene, grid [-{-scene},{-scene})^2 ({1 + 4 * scene * scene + 3} slots)" if scene < 0
            else f"scene {scene} final random spheres ({SLOTS[scene]} slots)")
    return f"{what}, {W}x{H}, {spp} spp, {depth} bounces"
FLOP_PER_TEST = 18          # 3 FADD + 3 FMUL + 6 FFMA per sphere test
SM_COUNT, FP32_LANES = 148, 128


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--split", default="rows", choices=["rows", "spp"])
    ap.add_argument("--spp-combine", default="gather", choices=["gather", "reduce"],
                    help="spp split: ordered gather (bit-exact) or NCCL sum-reduce of the accumulation buffer")
    ap.add_argument("--tile-rows", type=int, default=1)
    ap.add_argument("--accel", default="linear", choices=["linear", "lbvh", "grid"],
                    help="grid: experimental (csrc/rt_grid.cuh), needs RT_ENABLE_GRID=1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-lbvh-extra", action="store_true")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ------------------------------------------------------------------ clocks ------------------
class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------ CPU reference arm -------
def cpu_reference_step(width, height, spp, depth, procs):
    """One bounded sample of the workload on the reference's serial CPU renderer
    (oracle/_ref/inoneweekend_cpu = src/InOneWeekend compiled from its own headers), `procs`
    copies side by side (the program is single-threaded).  Returns (Mpath-samples/s aggregate,
    seconds, kind)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "inoneweekend_cpu")
    if os.path.exists(exe):
        t0 = time.perf_counter()
        ps = [subprocess.Popen([exe, str(width), str(height), str(spp), str(depth)],
                               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for _ in range(procs)]
        for p in ps:
            if p.wait() != 0:
                raise RuntimeError("inoneweekend_cpu failed")
        dt = time.perf_counter() - t0
        return procs * width * height * spp / dt / 1e6, dt, "reference"
    # the reference tree was not available at build time: time the oracle port instead
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from concurrent.futures import ThreadPoolExecutor
    slots, cam = O.scene(1), O.camera(width, height, spp, depth)
    bands = [(height * k // procs, height * (k + 1) // procs) for k in range(procs)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(procs) as ex:
        list(ex.map(lambda b: O.render(slots, cam, row0=b[0], row1=b[1]), bands))
    dt = time.perf_counter() - t0
    return width * height * spp / dt / 1e6, dt, "port"


CPU_SAMPLE = (480, 270, 8)      # width, height, spp of the bounded CPU sample (same scene/bounces)


def cpu_baseline(depth):
    procs = os.cpu_count() or 1
    w, h, spp = CPU_SAMPLE
    val, dt, kind = cpu_reference_step(w, h, spp, depth, procs)
    one, dt1, _ = cpu_reference_step(w, h, max(1, spp // 4), depth, 1)
    return {"value": round(val, 4), "unit": METRIC, "cores": procs, "kind": kind,
            "single_core_value": round(one, 4),
            "sample": f"scene 1, {w}x{h}, {spp} spp, {depth} bounces per process, {procs} concurrent "
                      f"processes of the serial src/InOneWeekend renderer ({dt:.1f} s)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene, W, H, spp, depth = WORKLOADS[args.workload]
    procs = os.cpu_count() or 1
    w, h, s = CPU_SAMPLE
    for _ in range(args.warmup):
        cpu_reference_step(w // 4, h // 4, 2, depth, procs)
    t0 = time.perf_counter()
    kind = "reference"
    for _ in range(args.steps):
        _, _, kind = cpu_reference_step(w, h, s, depth, procs)
    dt = time.perf_counter() - t0
    val = args.steps * procs * w * h * s / dt / 1e6
    sample = (f"scene 1, {w}x{h}, {s} spp, {depth} bounces per process x {procs} concurrent processes per step "
              f"(serial src/InOneWeekend renderer; bounded sample of {W}x{H}/{spp} spp)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": METRIC, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "implementation": "serial CPU src/InOneWeekend (double)",
                   "sample": sample},
        "cpu_baseline": {"value": round(val, 4), "unit": METRIC, "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": round(val, 4), "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------ reference kernel on the GPU
def reference_gpu_kernel():
    """The reference's kernels rebuilt for sm_100 (oracle/_ref), BASELINE config 2, their own render_ms stdout field:
    the global-float variant (the parity target) and, as timing comparators only, the constant- and texture-memory
    variants the shared-memory scene replaces (scene 1 only, ConstFloat main.cu:73-76)."""
    scene, W, H, spp, depth = WORKLOADS["cfg2"]

    def run(exe_name):
        exe = os.path.join(ROOT, "oracle", "_ref", exe_name)
        if not os.path.exists(exe):
            return None
        runs = []
        with tempfile.TemporaryDirectory() as tmp:
            for _ in range(2):
                out = subprocess.check_output([exe, "--scene_id", str(scene), "--width", str(W), "--height", str(H),
                                               "--samples", str(spp), "--bounces", str(depth), "--threads", "8"], cwd=tmp)
                runs.append(float(out.decode().split(",")[0]))
        ms = min(runs)
        return {"render_ms": round(ms, 3), "value": round(W * H * spp / ms / 1e3, 3)}

    g = run("global-float-cuda-raytrace")
    if g is None:
        return None
    line = {"kernel": "GlobalFloat render rebuilt -O3 sm_100, --threads 8", "workload": f"scene {scene}, {W}x{H}, {spp} spp, {depth} bounces",
            "render_ms": g["render_ms"], "value": g["value"], "unit": METRIC}
    for key, exe_name in (("const_float", "const-float-cuda-raytrace"), ("tex_float", "tex-float-cuda-raytrace")):
        try:
            v = run(exe_name)
        except Exception as e:                                    # a comparator that fails must not take the bench line with it
            v = {"error": str(e)[:120]}
        if v is not None:
            line[key] = v
    return line


# ------------------------------------------------------------------ B200 arm ----------------
class QuietStdout:
    """Route file descriptor 1 to stderr while the run is in progress: libraries (NCCL prints its version
    banner with printf) must not add lines to the one-JSON-line contract of stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    with QuietStdout():
        line = run_b200_arm(args)
    if line is not None:
        print(json.dumps(line), flush=True)


def run_b200_arm(args):

    import numpy as np
    import torch
    import torch.distributed as dist
    import raytracingincuda_b200 as rt
    from raytracingincuda_b200 import api
    from raytracingincuda_b200 import dist as rtdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    scene_id, W, H, spp, depth = WORKLOADS[args.workload]
    grid_accel = args.accel == "grid"
    lbvh = (scene_id < 0 or args.accel == "lbvh") and not grid_accel
    accel = api.ACCEL_GRID if grid_accel else (api.ACCEL_LBVH if lbvh else api.ACCEL_LINEAR)
    slots = rt.scene_scaled(-scene_id) if scene_id < 0 else rt.scene(scene_id)
    cam = rt.camera(W, H, spp, depth)
    chunks = rt.num_chunks(W, H, spp)
    r = rt.Renderer(local_rank)
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)
    r.upload_scene(slots)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    frame_dev = torch.empty((H, W, 3), dtype=torch.float32, device=device) if rank == 0 else None
    trace_ms, launches, segs, tests, nodes, binned = [], [0], [0], [0], [0], [0]

    def note_stats():
        st = r.stats()
        trace_ms.append(st.trace_ms)
        launches[0] += st.launches
        segs[0], tests[0], nodes[0], binned[0] = st.segments, st.sphere_tests, st.node_visits, st.binned_segments
        return st

    def step_device():
        """One render with everything resident on the device; result on rank 0's HBM."""
        if world == 1:
            r.render(cam, api.make_opts(accel=accel), out=frame_dev)
            note_stats()
            return frame_dev
        if args.split == "rows":
            def render_rows(buf):
                r.render(cam, api.make_opts(split=api.SPLIT_ROWS, rank=rank, world=world, tile_rows=args.tile_rows, accel=accel), out=buf)
                note_stats()
            return rtdist.render_rows_split(render_rows, W, H, args.tile_rows, rank, world, device, out=frame_dev)

        def render_partials(planes, c0, c1):
            r.render_partials(cam, api.make_opts(split=api.SPLIT_SPP, rank=rank, world=world, accel=accel), planes)
            note_stats()
        return rtdist.render_spp_split(render_partials, lambda planes: r.finalize(cam, planes, planes.shape[0], out=frame_dev),
                                       W, H, chunks, rank, world, device, combine=args.spp_combine)

    frame_host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory() if rank == 0 else None

    def step_e2e():
        """Public-API step with host buffers: scene slots from host memory in, frame in pinned
        host memory out (rank 0)."""
        r.upload_scene(slots)
        if world == 1:
            r.render(cam, api.make_opts(accel=accel), out=frame_host)
            note_stats()
            return
        out = step_device()
        if rank == 0:
            frame_host.copy_(out, non_blocking=True)
            torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(args.warmup):
        step_device()
    barrier()

    # ---- timed: device-resident ----
    trace_ms.clear()
    launches[0] = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            step_device()
        ev1.record(stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    timed_launches = launches[0]
    step_trace_ms = sum(trace_ms) / max(1, len(trace_ms))
    segments = segs[0]

    # ---- timed: end to end through the C ABI with host buffers ----
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([ms, e2e_s * 1e3, step_trace_ms], dtype=torch.float64, device=device)
    seg_t = torch.tensor([segments, tests[0], nodes[0], binned[0]], dtype=torch.int64, device=device)
    per_rank_ms = [step_trace_ms]
    if world > 1:
        # every rank's own path-tracing time per step: the step ends with the slowest rank
        mine = torch.tensor([step_trace_ms], dtype=torch.float64, device=device)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank_ms = [round(float(x.item()), 3) for x in allr]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(seg_t, op=dist.ReduceOp.SUM)
    ms, e2e_ms, step_trace_ms = [float(x) for x in t.tolist()]
    segments, sphere_tests, node_visits, binned_segments = [int(x) for x in seg_t.tolist()]

    line = None
    if rank == 0:
        peaks, peaks_src = measured_peaks()
        paths = W * H * spp
        ms_per_step = ms / args.steps
        value = paths / (ms_per_step * 1e-3) / 1e6
        e2e_value = paths / (e2e_ms / args.steps * 1e-3) / 1e6
        n_slots = len(slots)
        clk = clocks.summary()
        # linear scan: segments x slots tests; LBVH: the leaf/big tests the traversal actually made
        # (grid: the reference's algorithmic work as for the linear scan; the counted tests go into the note)
        flop = (segments * n_slots if grid_accel else sphere_tests) * FLOP_PER_TEST      # whole job, one step
        achieved = flop / (step_trace_ms * 1e-3) / 1e12 / world         # per GPU (per launch of the trace kernel)
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        peak = SM_COUNT * FP32_LANES * 2 * sm_max * 1e6 / 1e12
        obs = clk["sm_mhz"] or sm_max
        if lbvh:
            # LBVH: FP32 work actually asked for = 2 inflated slab tests per node visit (~52 FLOP: 18 sub/abs for the
            # distance bound and the slabs, 9 mul/fma, 4 for the inflation, 21 min/max/compare counted as 1 each) plus
            # 18 FLOP per exact sphere test.  The kernel is issue bound at 14-16 of 32 lanes active (ncu), not DRAM bound.
            flop_b = node_visits * 52 + sphere_tests * FLOP_PER_TEST
            ach = flop_b / (step_trace_ms * 1e-3) / 1e12 / world
            pk = SM_COUNT * FP32_LANES * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
            roof = {"bound": "fp32", "kernel": "trace_kernel<float,lbvh>", "unit": "TFLOP/s", "achieved": round(ach, 3),
                    "peak": round(pk, 2), "frac": round(ach / pk, 4),
                    "algorithmic": f"{node_visits} node visits x 52 FLOP + {sphere_tests} exact sphere tests x 18 FLOP per step",
                    "note": "divergent traversal: issue slots ~80 % busy with 14-16 of 32 lanes active; node records (64 B each, "
                            "6.4 MB for 99 860 slots) are L1/L2 resident",
                    "kernel_ms": round(step_trace_ms, 3), "traffic": None}
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": METRIC, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload),
                       "implementation": "float, " + ("uniform grid over the ground plane (EXPERIMENTAL, RT_ENABLE_GRID=1), camera rays "
                                                           "through per-tile candidate lists" if grid_accel else
                                                           "on-GPU LBVH" if lbvh else "linear scan in shared memory, camera rays "
                                                           "through per-tile candidate lists (rt_opts.primary_bins)"),
                       "l2": "inputs regenerate per step; "
                                   f"partial planes {chunks}x{W}x{H}x16 B exceed L2", "split": (args.split + ("/" + (args.spp_combine if args.split == "spp" else "nccl-gather"))) if world > 1 else "none",
                       "chunks": chunks, "seed": 1227},
            "render_ms": round(ms_per_step, 3),
            "e2e": {"value": round(e2e_value, 3), "unit": METRIC, "h2d_bytes_per_step": int(slots.nbytes),
                    "d2h_bytes_per_step": int(W * H * 3 * 4), "ms_per_step": round(e2e_ms / args.steps, 3)},
            "gpu_launches": timed_launches,
            "clocks": clk,
            "roofline": roof if lbvh else {"bound": "fp32", "kernel": "trace_kernel_pb<float,grid>" if grid_accel else "trace_kernel_pb<float>", "achieved": round(achieved, 3),
                         "peak": round(peak, 2), "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                         "frac_at_observed_clock": round(achieved / (peak * obs / sm_max), 4),
                         "peak_source": f"148 SM x 128 lanes x 2 x sm_max_mhz ({peaks_src} MEASURED_PEAKS.json)",
                         "algorithmic": (f"{sphere_tests} sphere tests x {FLOP_PER_TEST} FLOP per step ({node_visits} BVH node "
                                         f"visits not counted)" if lbvh else
                                         f"{segments} segments x {n_slots} slots x {FLOP_PER_TEST} FLOP per step"),
                         "note": ("achieved counts the reference's arithmetic (18 FLOP per (ray, slot) test of every segment, SURVEY 8d); "
                                  f"the kernel resolves the {binned_segments} camera-ray segments against per-tile candidate lists "
                                  f"and scans the other {segments - binned_segments} with a 7-FMA conservative filter per test plus "
                                  "the exact test on the ~0.5 % candidates"),
                         "scanned_fraction": round((segments - binned_segments) / max(1, segments), 4),
                         **({"grid": {"cells_per_segment": round(node_visits / max(1, segments), 3),
                                      "exact_tests_per_segment": round(sphere_tests / max(1, segments), 3)}} if grid_accel else {}),
                         "kernel_ms": round(step_trace_ms, 3), "traffic": None},
        }
        # DRAM bytes per trace_kernel launch from the committed ncu pass over this same command
        # (profiles/r01e_bench_kernel_traffic.json); None for workloads that were not captured
        try:
            with open(os.path.join(ROOT, "profiles", "r01e_bench_kernel_traffic.json")) as f:
                cap = json.load(f).get(args.workload, {})
            k = next(v for name, v in cap.items() if "trace_kernel" in name)
            if world == 1 and not lbvh:
                line["roofline"]["traffic"] = int(k["dram_read_bytes_per_launch"] + k["dram_write_bytes_per_launch"])
                line["roofline"]["traffic_note"] = ("ncu dram__bytes_read+write per launch; algorithmic HBM bytes are the "
                                                    f"{chunks}x{W}x{H}x16 B partial planes written once = {chunks * W * H * 16} B")
        except Exception:
            pass
        st = r.stats()
        if world > 1:
            line["per_rank_kernel_ms"] = per_rank_ms
        line["kernel"] = {"grid": st.grid, "block": st.block, "regs": st.regs, "smem_bytes": st.smem_bytes,
                          "segments_per_path": round(segments / paths, 4),
                          "binned_segments_per_path": round(binned_segments / paths, 4)}
        if world == 1 and not lbvh and not args.no_lbvh_extra:
            # the same workload with every segment through the shared-memory scan (rt_opts.primary_bins off): same image
            oo = api.make_opts(primary_bins=api.PBINS_OFF)
            r.render(cam, oo, out=frame_dev)
            r.render(cam, oo, out=frame_dev)
            so = r.stats()
            line["scan_every_segment"] = {"value": round(paths / (so.render_ms * 1e-3) / 1e6, 3), "unit": METRIC,
                                          "ms_per_step": round(so.render_ms, 3), "kernel": "trace_kernel<float,linear>",
                                          "roofline_frac": round(so.segments * n_slots * FLOP_PER_TEST / (so.trace_ms * 1e-3) / 1e12 / peak, 4)}
        if world == 1 and not lbvh and n_slots >= 256 and not args.no_lbvh_extra:
            # the same workload through the on-GPU LBVH (bit-identical image): informative, not the headline
            ob = api.make_opts(accel=api.ACCEL_LBVH)
            r.render(cam, ob, out=frame_dev)
            t_l = []
            for _ in range(2):
                r.render(cam, ob, out=frame_dev)
                t_l.append(r.stats().render_ms)
            sb = r.stats()
            line["accel_lbvh"] = {"value": round(paths / (min(t_l) * 1e-3) / 1e6, 3), "unit": METRIC,
                                  "ms_per_step": round(min(t_l), 3), "node_visits_per_segment": round(sb.node_visits / sb.segments, 2),
                                  "sphere_tests_per_segment": round(sb.sphere_tests / sb.segments, 2)}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(depth)
        if world == 1 and not args.no_ref_gpu:
            try:
                line["reference_gpu_kernel"] = reference_gpu_kernel()
            except Exception as e:  # the comparator is informative, never fatal
                line["reference_gpu_kernel"] = {"error": str(e)[:200]}
    r.close()
    if world > 1:
        dist.destroy_process_group()
    return line if rank == 0 else None


if __name__ == "__main__":
    main()
