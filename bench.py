#!/usr/bin/env python
"""bench.py -- headline benchmark of the render hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full render of the workload (default: BASELINE config 4, scene 1, 3840x2160,
1000 spp, 50 bounces, float) through the library's DEFAULT options (rt_opts_default: RT_ACCEL_AUTO --
what a user of the drop-in gets).  `value` = W*H*spp*1e-6 / step-seconds (Mpath-samples/s), timed with
CUDA events on the launching stream, scene already resident in HBM, max over ranks.  `e2e` is the
same metric through the public C-ABI call with HOST buffers: per step the scene slots are uploaded
from host memory and the gamma-encoded frame is read back into pinned host memory.
Besides the headline the same JSON line carries (N = 1) every other BASELINE config (`configs`: config 2,
config 3 scenes 2 and 3 in float and double, config 5), the shared-memory linear scan on the headline workload
with its executed-FP32 roofline (`linear_scan`: the kernel the north star's FP32-pipe target is about), the CPU
baseline, the reference's own GPU kernels rebuilt for sm_100 and the CLI's stdout contract at 4K; and (N > 1) both
multi-GPU partitionings (`splits`) and config 5 on N GPUs.
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
FLOP_PER_TEST = 18          # 3 FADD + 3 FMUL + 6 FFMA per exact sphere test (the reference's arithmetic)
FLOP_PER_FILTER = 14        # the scan's conservative filter: 7 FMA per (ray, slot) test
FLOP_PER_NODE = {"lbvh": 52, "grid": 24}   # two inflated slab tests per LBVH node visit; cell exit + index arithmetic per grid cell
SM_COUNT, FP32_LANES = 148, 128


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--split", default="spp", choices=["rows", "spp"],
                    help="headline partitioning for --gpus > 1 (both are timed and reported in `splits`; spp is balanced by construction, "
                         "rows depends on how evenly the cost of a frame is spread over its rows)")
    ap.add_argument("--tile-rows", type=int, default=1)
    ap.add_argument("--accel", default="auto", choices=["auto", "linear", "lbvh", "grid"],
                    help="auto (default) = rt_opts_default: what the drop-in binary uses")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only: no configs / linear_scan / splits blocks")
    ap.add_argument("--no-lbvh-extra", action="store_true", help=argparse.SUPPRESS)      # old name of --no-extras
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


# ------------------------------------------------------------------ reference kernels on the GPU
def _run_ref(exe_name, scene, W, H, spp, depth, runs):
    """render_ms values (the reference's own stdout field, GF main.cu:342-343) of `runs` runs of a rebuilt reference binary."""
    exe = os.path.join(ROOT, "oracle", "_ref", exe_name)
    if not os.path.exists(exe):
        return None
    out_ms, e2e = [], []
    with tempfile.TemporaryDirectory() as tmp:
        for _ in range(runs):
            out = subprocess.check_output([exe, "--scene_id", str(scene), "--width", str(W), "--height", str(H),
                                           "--samples", str(spp), "--bounces", str(depth), "--threads", "8"], cwd=tmp)
            a, b = out.decode().split(",")
            out_ms.append(float(a))
            e2e.append(float(b))
    ms = statistics.median(out_ms)
    return {"render_ms": round(ms, 3), "render_ms_min": round(min(out_ms), 3), "e2e_ms": round(statistics.median(e2e), 3),
            "runs": runs, "value": round(W * H * spp / ms / 1e3, 3)}


def reference_gpu_kernels(full):
    """The reference's own kernels rebuilt -O3 for sm_100 (oracle/_ref), timed by their own render_ms stdout field, --threads 8:
    GlobalFloat at config 2 (median of 3, SURVEY 8d comparator (ii)) and, when `full`, once at the headline config 4 and
    GlobalDouble at config 3; the constant- and texture-memory variants at config 2 as timing comparators only."""
    scene, W, H, spp, depth = WORKLOADS["cfg2"]
    g = _run_ref("global-float-cuda-raytrace", scene, W, H, spp, depth, 3)
    if g is None:
        return None
    line = {"kernel": "GlobalFloat render rebuilt -O3 sm_100, --threads 8", "unit": METRIC,
            "cfg2": dict(g, workload=workload_name("cfg2"))}
    line.update({"workload": workload_name("cfg2"), "render_ms": g["render_ms"], "value": g["value"]})
    extra = [("const_float", "const-float-cuda-raytrace", "cfg2", 1), ("tex_float", "tex-float-cuda-raytrace", "cfg2", 1)]
    if full:
        extra += [("cfg4", "global-float-cuda-raytrace", "cfg4", 1), ("cfg3a_double", "global-double-cuda-raytrace", "cfg3a", 1),
                  ("cfg3b_double", "global-double-cuda-raytrace", "cfg3b", 1), ("cfg3a", "global-float-cuda-raytrace", "cfg3a", 1),
                  ("cfg3b", "global-float-cuda-raytrace", "cfg3b", 1)]
    for key, exe_name, wl, runs in extra:
        try:
            v = _run_ref(exe_name, *WORKLOADS[wl], runs)
        except Exception as e:                                    # a comparator that fails must not take the bench line with it
            v = {"error": str(e)[:120]}
        if v is not None:
            line[key] = dict(v, workload=workload_name(wl))
    return line


def cli_contract():
    """The drop-in binary's own stdout contract (`render_ms,e2e_ms`, GF main.cu:342-343,397-398) at the headline config, PPM
    write included (8.3 M text lines): one run, default options."""
    exe = os.path.join(ROOT, "raytracingincuda_b200", "bin", "b200-raytrace")
    scene, W, H, spp, depth = WORKLOADS["cfg4"]
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        out = subprocess.check_output([exe, "--scene_id", str(scene), "--width", str(W), "--height", str(H), "--samples", str(spp),
                                       "--bounces", str(depth), "--threads", "8"], cwd=tmp)
        wall = time.perf_counter() - t0
        ppm = [f for f in os.listdir(tmp) if f.endswith(".ppm")]
        size = os.path.getsize(os.path.join(tmp, ppm[0])) if ppm else 0
    a, b = out.decode().split(",")
    return {"workload": workload_name("cfg4"), "render_ms": round(float(a), 3), "e2e_ms": round(float(b), 3),
            "process_wall_ms": round(wall * 1e3, 1), "ppm_bytes": size,
            "note": "e2e_ms: scene build + upload + render + D2H + P3 PPM write, CUDA context creation excluded (GF main.cu:81-95,394-400)"}


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


def fp32_peak(peaks):
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    return SM_COUNT * FP32_LANES * 2 * sm_max * 1e6 / 1e12


def lib_id():
    """Hash of the device sources the loaded library was built from (ncu captures under profiles/ carry the same id)."""
    import raytracingincuda_b200.api as api
    return api.lib().rt_kernel_build_id().decode()


def ncu_capture(kind):
    """Pipe / issue utilisation of a trace kernel from the committed ncu capture (profiles/r02_pb_<kind>_key_metrics.csv), only
    when the capture was taken with the library build that is loaded now (profiles/r02_capture_id.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_capture_id.json")) as f:
            cap_id = json.load(f)["lib_id"]
        if cap_id != lib_id():
            return {"stale": f"profiles/r02_pb_{kind}_* were captured with build {cap_id}, this is {lib_id()}"}
        out = {"source": f"profiles/r02_pb_{kind}_key_metrics.csv (ncu --set full, 1920x1080/16 spp, same library build)"}
        names = {"sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_cycles_active_pct",
                 "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
                 "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_instruction",
                 "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
                 "launch__registers_per_thread": "registers"}
        with open(os.path.join(ROOT, "profiles", f"r02_pb_{kind}_key_metrics.csv")) as f:
            for row in f:
                k, _, v = row.strip().split(",")
                if k in names:
                    out[names[k]] = round(float(v), 2)
        return out
    except Exception as e:
        return {"unavailable": str(e)[:80]}


def framebuffer_roofline(r, job, peaks, peaks_src, reps=20):
    """finalize_flat_kernel on the job's frame: 24 B of int64 accumulators read + 12 B of float frame written per pixel."""
    import torch
    from raytracingincuda_b200 import api
    import raytracingincuda_b200 as rt
    cam1 = rt.camera(job.W, job.H, 1, job.depth)
    acc = torch.empty((job.W * job.H * 3,), dtype=torch.int64, device=job.frame.device)
    r.render_partials(cam1, api.make_opts(), acc)              # real radiance sums of one sample per pixel
    ms = []
    for _ in range(reps + 3):
        r.finalize(cam1, acc, out=job.frame)
        ms.append(r.last_finalize_ms)
    med = statistics.median(ms[3:])
    os.environ["RT_FINALIZE_BY_PIXEL"] = "1"                   # the four-pixels-per-thread kernel this one replaced (still used for row placement)
    try:
        old = []
        for _ in range(8):
            r.finalize(cam1, acc, out=job.frame)
            old.append(r.last_finalize_ms)
    finally:
        del os.environ["RT_FINALIZE_BY_PIXEL"]
    nbytes = job.W * job.H * (24 + 12)
    peak = float(peaks.get("hbm_gbs", 6650.0))
    out = {"bound": "hbm", "kernel": "finalize_flat_kernel<float>", "unit": "GB/s", "bytes_per_launch": nbytes, "ms": round(med, 4),
           "ms_min": round(min(ms[3:]), 4), "launches_timed": reps, "achieved": round(nbytes / (med * 1e-3) / 1e9, 1), "peak": peak,
           "frac": round(nbytes / (med * 1e-3) / 1e9 / peak, 4), "peak_source": f"{peaks_src} MEASURED_PEAKS.json hbm_gbs (copy bandwidth)",
           "l2": f"{job.W * job.H * 24} B read + {job.W * job.H * 12} B written per launch: " +
                 ("larger than the 126 MB L2" if nbytes > 126e6 else "fits the 126 MB L2 (not an HBM measurement at this frame size)"), "traffic": None,
           "by_pixel_kernel_ms": round(statistics.median(old[3:]), 4)}
    try:
        with open(os.path.join(ROOT, "profiles", "r02_bench_kernel_traffic.json")) as f:
            cap = json.load(f)
        if cap.get("lib_id") == lib_id():
            k = next((v for name, v in cap.get("cfg4", {}).items() if "finalize" in name), None)
            if k and (job.W, job.H) == (3840, 2160):
                out["traffic"] = int(k["dram_read_bytes_per_launch"] + k["dram_write_bytes_per_launch"])
                out["traffic_source"] = "profiles/r02_bench_kernel_traffic.json (ncu dram__bytes_read+write per launch, same library build)"
    except Exception:
        pass
    return out


def work_of(st, accel_name, n_slots, ms, peak):
    """Executed and reference-equivalent FP32 work of one trace launch from the kernel's own counters (rt_stats)."""
    executed = st.filter_tests * FLOP_PER_FILTER + st.sphere_tests * FLOP_PER_TEST + st.node_visits * FLOP_PER_NODE.get(accel_name, 0)
    ref_equiv = st.segments * n_slots * FLOP_PER_TEST
    sec = ms * 1e-3
    return {"executed_tflops": round(executed / sec / 1e12, 3), "frac_executed": round(executed / sec / 1e12 / peak, 4),
            "reference_equivalent_tflops": round(ref_equiv / sec / 1e12, 3),
            "filter_tests": int(st.filter_tests), "exact_sphere_tests": int(st.sphere_tests), "node_visits": int(st.node_visits),
            "segments": int(st.segments), "binned_segments": int(st.binned_segments)}


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
    extras = not (args.no_extras or args.no_lbvh_extra)
    peaks, peaks_src = measured_peaks()
    peak = fp32_peak(peaks)
    ACCEL = {"auto": api.ACCEL_AUTO, "linear": api.ACCEL_LINEAR, "lbvh": api.ACCEL_LBVH, "grid": api.ACCEL_GRID}

    r = rt.Renderer(local_rank)
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def scene_of(scene_id, double=False):
        return rt.scene_scaled(-scene_id) if scene_id < 0 else rt.scene(scene_id, double=double)

    # ------------------------------------------------------------------ one workload on this world -----------------
    class Job:
        """A workload bound to the renderer: device-resident steps for either partitioning."""

        def __init__(self, name, accel="auto", double=False):
            self.name, self.accel_name, self.double = name, accel, double
            self.scene_id, self.W, self.H, self.spp, self.depth = WORKLOADS[name]
            self.slots = scene_of(self.scene_id, double)
            self.cam = rt.camera(self.W, self.H, self.spp, self.depth, double=double)
            self.dtype = torch.float64 if double else torch.float32
            self.frame = torch.empty((self.H, self.W, 3), dtype=self.dtype, device=device) if rank == 0 else None
            self.acc = None
            self.trace_ms, self.launches, self.st = [], 0, None
            self.paths = self.W * self.H * self.spp

        def upload(self):
            r.upload_scene(self.slots)

        def note(self):
            self.st = r.stats()
            self.trace_ms.append(self.st.trace_ms)
            self.launches += self.st.launches

        def opts(self, **kw):
            return api.make_opts(accel=ACCEL[self.accel_name], **kw)

        def step(self, split="rows", out=None):
            out = self.frame if out is None else out
            if world == 1:
                r.render(self.cam, self.opts(), out=out)
                self.note()
                return out
            if split == "rows":
                def render_rows(buf):
                    r.render(self.cam, self.opts(split=api.SPLIT_ROWS, rank=rank, world=world, tile_rows=args.tile_rows), out=buf)
                    self.note()
                return rtdist.render_rows_split(render_rows, self.W, self.H, args.tile_rows, rank, world, device, out=out)

            def render_partials(acc, s0, s1):
                r.render_partials(self.cam, self.opts(split=api.SPLIT_SPP, rank=rank, world=world), acc)
                self.note()
            if self.acc is None:
                self.acc = torch.empty((self.H, self.W, 3), dtype=torch.int64, device=device)

            def fin(acc):
                r.finalize(self.cam, acc, out=out)
                self.launches += 1
                return out
            return rtdist.render_spp_split(render_partials, fin, self.W, self.H, self.spp, rank, world, device, acc=self.acc)

        def timed(self, steps, warmup, split="rows"):
            """(ms per step over all ranks, per-rank trace ms, stats of the last launch) -- CUDA events, max over ranks."""
            for _ in range(warmup):
                self.step(split)
            barrier()
            self.trace_ms.clear()
            self.launches = 0
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(steps):
                self.step(split)
            ev1.record(stream)
            barrier()
            ms = ev0.elapsed_time(ev1) / steps
            mine = sum(self.trace_ms) / max(1, len(self.trace_ms))
            t = torch.tensor([ms, mine], dtype=torch.float64, device=device)
            per_rank = [round(mine, 3)]
            counters = torch.tensor([self.st.segments, self.st.sphere_tests, self.st.node_visits, self.st.binned_segments,
                                     self.st.filter_tests, self.st.paths], dtype=torch.int64, device=device)
            if world > 1:
                allr = [torch.zeros(1, dtype=torch.float64, device=device) for _ in range(world)]
                dist.all_gather(allr, t[1:2].clone())
                per_rank = [round(float(x.item()), 3) for x in allr]
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(counters, op=dist.ReduceOp.SUM)
            return float(t[0]), float(t[1]), per_rank, [int(x) for x in counters.tolist()]

    def summary(job, ms, kernel_ms, counters):
        """value + executed-work roofline of one measured workload (whole job, all ranks)."""
        seg, exact, nodes, binned, filt, paths = counters
        used = api.ACCEL_NAMES.get(job.st.accel_used, "?")
        executed = filt * FLOP_PER_FILTER + exact * FLOP_PER_TEST + nodes * FLOP_PER_NODE.get(used, 0)
        ach = executed / (kernel_ms * 1e-3) / 1e12 / world
        return {"workload": workload_name(job.name), "dtype": "f64" if job.double else "f32", "accel": used,
                "value": round(job.paths / (ms * 1e-3) / 1e6, 3), "unit": METRIC, "ms_per_step": round(ms, 3),
                "kernel_ms": round(kernel_ms, 3), "segments_per_path": round(seg / max(1, paths), 4),
                "executed_tflops_per_gpu": round(ach, 3), "frac_executed": round(ach / peak, 4),
                "exact_tests_per_segment": round(exact / max(1, seg), 3), "filter_tests_per_segment": round(filt / max(1, seg), 2),
                "node_visits_per_segment": round(nodes / max(1, seg), 3), "binned_segments_per_path": round(binned / max(1, paths), 4),
                "slots": len(job.slots), "regs": job.st.regs, "grid": job.st.grid}

    # ------------------------------------------------------------------ headline -----------------------------------
    job = Job(args.workload, args.accel)
    job.upload()
    frame_host = torch.empty((job.H, job.W, 3), dtype=torch.float32).pin_memory() if rank == 0 else None
    with ClockSampler(local_rank) as clocks:
        ms, kernel_ms, per_rank_ms, counters = job.timed(args.steps, args.warmup, args.split)
    timed_launches = job.launches
    head = summary(job, ms, kernel_ms, counters)
    build_ms = {"bvh_build_ms": round(job.st.bvh_build_ms, 3), "grid_build_ms": round(job.st.grid_build_ms, 3)}

    # ---- end to end through the C ABI with host buffers: scene slots from host memory in, frame in pinned host memory out ----
    def step_e2e():
        r.upload_scene(job.slots)
        if world == 1:
            r.render(job.cam, job.opts(), out=frame_host)
            return
        out = job.step(args.split)
        if rank == 0:
            frame_host.copy_(out, non_blocking=True)
            torch.cuda.synchronize()
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e[0])

    line = None
    if rank == 0:
        clk = clocks.summary()
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        obs = clk["sm_mhz"] or sm_max
        used = head["accel"]
        kernel_name = {"linear": "trace_kernel_pb<float,linear>", "lbvh": "trace_kernel_pb<float,lbvh>", "grid": "trace_kernel_pb<float,grid>"}.get(used, used)
        acc_bytes = job.W * job.H * 24
        roof = {"bound": "fp32", "kernel": kernel_name, "unit": "TFLOP/s",
                "achieved": head["executed_tflops_per_gpu"], "peak": round(peak, 2), "frac": head["frac_executed"],
                "frac_at_observed_clock": round(head["executed_tflops_per_gpu"] / (peak * obs / sm_max), 4),
                "peak_source": f"148 SM x 128 lanes x 2 x sm_max_mhz ({peaks_src} MEASURED_PEAKS.json)",
                "algorithmic": (f"EXECUTED FP32 work per step from the kernel's counters: {counters[4]} filter tests x {FLOP_PER_FILTER} + "
                                f"{counters[1]} exact sphere tests x {FLOP_PER_TEST} + {counters[2]} node/cell visits x "
                                f"{FLOP_PER_NODE.get(used, 0)} FLOP"),
                "reference_equivalent_tflops": round(counters[0] * head["slots"] * FLOP_PER_TEST / (kernel_ms * 1e-3) / 1e12 / world, 3),
                "reference_equivalent_note": "SURVEY 8d formula (segments x slots x 18 FLOP): the reference's work, not this kernel's; "
                                             "it exceeds the peak as soon as an algorithm culls tests, so it is not a roofline fraction",
                "kernel_ms": round(kernel_ms, 3),
                "note": ("the traversal kernels (grid / LBVH) are bound by divergence and dependent-load latency, not by the FP32 pipe; the "
                         "FP32-bound kernel of this path is the shared-memory linear scan: see `linear_scan`") if used != "linear" else
                        "FP32 issue bound: 7 FFMA2 + LDS.128 + 2 FSETP + 2 predicated OR per record of two tests",
                "ncu": ncu_capture({"linear": "linear", "grid": "grid", "lbvh": "lbvh_s1"}.get(used, used)),
                "traffic": None}
        # HBM traffic per trace launch: from the committed ncu pass of THIS library build when there is one, else what the
        # launch is known to move (accumulators zeroed, read-modify-written band by band through L2, read by finalize)
        roof["traffic_algorithmic"] = int(acc_bytes * 3 + job.W * job.H * 12)
        roof["traffic_algorithmic_note"] = (f"{job.W}x{job.H} pixels x (24 B accumulators: zeroed + written back once + read by finalize_kernel) + 12 B frame "
                                            "store; the round-1 build moved 32 partial planes = 8.5 GB per frame")
        try:
            with open(os.path.join(ROOT, "profiles", "r02_bench_kernel_traffic.json")) as f:
                cap = json.load(f)
            if cap.get("lib_id") == lib_id():
                k = next((v for name, v in cap.get(args.workload, {}).items() if "trace_kernel" in name), None)
                if k and world == 1:
                    roof["traffic"] = int(k["dram_read_bytes_per_launch"] + k["dram_write_bytes_per_launch"])
                    roof["traffic_source"] = "profiles/r02_bench_kernel_traffic.json (ncu dram__bytes_read+write per launch, same library build)"
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": head["value"], "unit": METRIC, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload),
                       "implementation": f"float, rt_opts_default (accel {args.accel} -> {used}), camera rays through per-tile candidate lists, "
                                         "64-bit fixed-point accumulation with integer atomics",
                       "l2": f"inputs regenerate per step; the accumulators ({acc_bytes} B) and the frame exceed L2 at 4K and are rewritten every step",
                       "split": (args.split + ("/nccl-int64-reduce" if args.split == "spp" else "/nccl-gather")) if world > 1 else "none",
                       "jobs_per_pixel": job.st.chunks, "seed": 1227, "lib_id": lib_id()},
            "render_ms": head["ms_per_step"],
            "e2e": {"value": round(job.paths / (e2e_ms * 1e-3) / 1e6, 3), "unit": METRIC, "h2d_bytes_per_step": int(job.slots.nbytes),
                    "d2h_bytes_per_step": int(job.W * job.H * 3 * 4), "ms_per_step": round(e2e_ms, 3)},
            "gpu_launches": timed_launches,
            "clocks": clk,
            "roofline": roof,
            "kernel": dict({k: head[k] for k in ("accel", "regs", "grid", "segments_per_path", "binned_segments_per_path",
                                                 "exact_tests_per_segment", "filter_tests_per_segment", "node_visits_per_segment")},
                           block=job.st.block, smem_bytes=job.st.smem_bytes, **build_ms),
        }
        if world > 1:
            line["per_rank_kernel_ms"] = per_rank_ms

    # ------------------------------------------------------------------ extras (all ranks walk the same sequence) ---
    if extras and world > 1:
        splits = {}
        for sp in ("rows", "spp"):
            if sp == args.split:
                if rank == 0:
                    splits[sp] = {"value": head["value"], "ms_per_step": head["ms_per_step"], "per_rank_kernel_ms": per_rank_ms, "headline": True}
                continue
            m, km, pr, _ = job.timed(max(2, min(5, args.steps)), 1, sp)
            if rank == 0:
                splits[sp] = {"value": round(job.paths / (m * 1e-3) / 1e6, 3), "ms_per_step": round(m, 3), "per_rank_kernel_ms": pr,
                              "exchange": "one NCCL sum-reduce of the int64 accumulation buffer (W*H*24 B per rank), bit-identical to 1 GPU"
                              if sp == "spp" else "NCCL gather of each rank's finished rows"}
        if rank == 0:
            line["splits"] = splits
        if args.workload == "cfg4":
            j5 = Job("cfg5", "lbvh")         # the structure BASELINE config 5 names; the library's default for this field is the grid
            j5.upload()
            m, km, pr, c5 = j5.timed(2, 1, "spp")
            m_r, km_r, pr_r, _ = j5.timed(2, 1, "rows")
            j5g = Job("cfg5", "auto")
            j5g.upload()
            m_g, km_g, pr_g, c5g = j5g.timed(2, 1, "spp")
            if rank == 0:
                line["cfg5"] = dict(summary(j5, m, km, c5), per_rank_kernel_ms=pr, bvh_build_ms=round(j5.st.bvh_build_ms, 3), split="spp",
                                    rows_split={"value": round(j5.paths / (m_r * 1e-3) / 1e6, 3), "ms_per_step": round(m_r, 3), "per_rank_kernel_ms": pr_r,
                                                "note": "rows of the horizon cost many times the average (camera rays of overflowing tiles traverse the whole "
                                                        "field), so a rank's share of the time depends on which rows it owns"},
                                    default_path=dict(summary(j5g, m_g, km_g, c5g), per_rank_kernel_ms=pr_g, split="spp"))
            job.upload()
    if extras and world == 1 and rank == 0:
        def quick(name, accel="auto", double=False, reps=2, warm=1):
            j = Job(name, accel, double)
            j.upload()
            m, km, _, c = j.timed(reps, warm)
            d = summary(j, m, km, c)
            d["bvh_build_ms"], d["grid_build_ms"] = round(j.st.bvh_build_ms, 3), round(j.st.grid_build_ms, 3)
            return d
        # the FP32-bound kernel of this path on the headline workload: the roofline exhibit
        lin = quick(args.workload, "linear") if head["accel"] != "linear" or args.accel != "linear" else dict(head)
        lin["roofline"] = {"bound": "fp32", "kernel": "trace_kernel_pb<float,linear>", "unit": "TFLOP/s", "achieved": lin["executed_tflops_per_gpu"],
                           "peak": round(peak, 2), "frac": lin["frac_executed"],
                           "reference_equivalent_tflops": round(lin["segments_per_path"] * job.paths * lin["slots"] * FLOP_PER_TEST
                                                                / (lin["kernel_ms"] * 1e-3) / 1e12, 3),
                           "ncu": ncu_capture("linear")}
        line["linear_scan"] = lin
        if len(job.slots) >= 256:
            line["accel_lbvh"] = quick(args.workload, "lbvh")
            if head["accel"] != "grid":
                try:
                    line["accel_grid"] = quick(args.workload, "grid")
                except Exception as e:
                    line["accel_grid"] = {"error": str(e)[:100]}
        # the one HBM-bound kernel of the path: accumulators -> gamma-encoded frame (north star (d), "achieved HBM GB/s for the
        # framebuffer"); timed live through rt_finalize on a buffer of the headline frame's size, input + output larger than L2
        try:
            line["framebuffer"] = framebuffer_roofline(r, job, peaks, peaks_src)
        except Exception as e:
            line["framebuffer"] = {"error": str(e)[:120]}
        if args.workload == "cfg4":
            cfgs = {}
            # config 5 names its structure ("on-GPU LBVH"): it is timed through RT_ACCEL_LBVH; what the library's default
            # (RT_ACCEL_AUTO: the uniform grid on this field since it overtook the LBVH) makes of the same frame sits beside it
            for key, name, dbl, acc in (("cfg2", "cfg2", False, "auto"), ("cfg3_scene2", "cfg3a", False, "auto"), ("cfg3_scene3", "cfg3b", False, "auto"),
                                        ("cfg3_scene2_double", "cfg3a", True, "auto"), ("cfg3_scene3_double", "cfg3b", True, "auto"),
                                        ("cfg5", "cfg5", False, "lbvh"), ("cfg5_default_path", "cfg5", False, "auto")):
                try:
                    cfgs[key] = quick(name, acc, dbl, reps=3)
                except Exception as e:
                    cfgs[key] = {"error": str(e)[:100]}
            line["configs"] = cfgs
        job.upload()
    if world == 1 and rank == 0:
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(job.depth)
        if not args.no_ref_gpu:
            try:
                line["reference_gpu_kernel"] = reference_gpu_kernels(full=extras and args.workload == "cfg4")
            except Exception as e:  # the comparator is informative, never fatal
                line["reference_gpu_kernel"] = {"error": str(e)[:200]}
            if extras and args.workload == "cfg4":
                try:
                    line["cli"] = cli_contract()
                except Exception as e:
                    line["cli"] = {"error": str(e)[:200]}
    r.close()
    if world > 1:
        dist.destroy_process_group()
    return line if rank == 0 else None


if __name__ == "__main__":
    main()
