#!/usr/bin/env python
"""Benchmark sweep with the reference's CSV schema (global_float_benchmark.sh:25-82) and the
averaging step of timing-benchmarks/process.py:16-33 in one tool, so rows of the new binary and of
the reference rebuilt for sm_100 land in tables that can be joined.

    python tools/benchmark.py --exe raytracingincuda_b200/bin/b200-raytrace --out benchmarks/b200.csv \
        --scenes 1 --sizes 320x192,1280x768,1920x1080 --samples 100 --bounces 25 --threads 8 --runs 5 [-- extra flags]

Writes <out> (one row per run: scene_id,width,height,samples,bounces,threads,run,render_only_time_ms,
end_to_end_time_ms) and <out minus .csv>_avg.csv (group means + Mpath-samples/s).
"""
import argparse
import csv
import itertools
import os
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    argv = sys.argv[1:]
    extra = []
    if "--" in argv:
        k = argv.index("--")
        argv, extra = argv[:k], argv[k + 1:]
    ap = argparse.ArgumentParser()
    ap.add_argument("--exe", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--scenes", default="1")
    ap.add_argument("--sizes", default="320x192,480x288,640x384,960x576,1280x768")      # the reference's sweep
    ap.add_argument("--samples", default="100")
    ap.add_argument("--bounces", default="25")
    ap.add_argument("--threads", default="8")
    ap.add_argument("--runs", type=int, default=5)
    a = ap.parse_args(argv)
    ints = lambda s: [int(x) for x in s.split(",")]
    sizes = [tuple(int(v) for v in s.split("x")) for s in a.sizes.split(",")]
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    exe = os.path.abspath(a.exe)
    rows = []
    with open(a.out, "w", newline="") as f, tempfile.TemporaryDirectory() as cwd:
        w = csv.writer(f)
        w.writerow(["scene_id", "width", "height", "samples", "bounces", "threads", "run",
                    "render_only_time_ms", "end_to_end_time_ms"])
        for threads, scene, samples, bounces, (width, height) in itertools.product(
                ints(a.threads), ints(a.scenes), ints(a.samples), ints(a.bounces), sizes):
            for run in range(1, a.runs + 1):
                cmd = [exe, "--scene_id", str(scene), "--width", str(width), "--height", str(height),
                       "--samples", str(samples), "--bounces", str(bounces), "--threads", str(threads)] + extra
                p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
                fields = [x.strip() for x in p.stdout.strip().split(",")] if p.returncode == 0 else ["", ""]
                if len(fields) != 2:
                    fields = ["", ""]                      # the reference leaves the cells empty on failure
                w.writerow([scene, width, height, samples, bounces, threads, run] + fields)
                f.flush()
                rows.append(((scene, width, height, samples, bounces, threads), fields))
    groups = defaultdict(list)
    for key, fields in rows:
        if fields[0]:
            groups[key].append((float(fields[0]), float(fields[1])))
    avg_path = a.out[:-4] + "_avg.csv" if a.out.endswith(".csv") else a.out + "_avg.csv"
    with open(avg_path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["scene_id", "width", "height", "samples", "bounces", "threads", "avg_render_only_time_ms",
                    "avg_end_to_end_time_ms", "mpath_samples_per_s"])
        for key in sorted(groups):
            r = sum(x[0] for x in groups[key]) / len(groups[key])
            e = sum(x[1] for x in groups[key]) / len(groups[key])
            w.writerow(list(key) + [f"{r:.6f}", f"{e:.6f}", f"{key[1] * key[2] * key[3] / r / 1e3:.3f}"])
    print(f"wrote {a.out} and {avg_path}")


if __name__ == "__main__":
    main()
