#!/usr/bin/env python
"""How many steps would a uniform grid over the ground plane need per ray segment?  (CPU analysis, no GPU.)

The reference's scenes are fields of equal small spheres standing on a plane, one per unit cell, plus a few big ones.
The LBVH needs 6.2 (scene 1) / 13.2 (99 860 slots) node visits per segment with the camera rays binned
(profiles/README.md).  This script logs real path segments with the CPU oracle (orc_log_segments), and counts for every
SCATTERED segment the cells a 2-D DDA over (x, z) would visit: the ray is clipped to the slab that holds the small
spheres, to the grid's bounds and to its hit distance; cells = 1 + |floor dx| + |floor dz| between entry and exit.
It prints the distribution per ray and per group of 32 rays (the lock-step cost of a warp is its slowest lane).

usage: python tools/analyse_accel.py [paths]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import oracle_lib as O  # noqa: E402


def segments_of(slots, cam, n_paths, rng):
    L = O.lib()
    L.orc_log_segments.argtypes = [C.c_void_p, C.c_long]
    L.orc_logged_segments.restype = C.c_long
    buf = np.zeros((n_paths * 60, 9), dtype=np.float32)
    L.orc_log_segments(buf.ctypes.data, len(buf))
    rgb = np.zeros(3, dtype=np.float32)
    for _ in range(n_paths):
        i, j = int(rng.integers(0, cam.width)), int(rng.integers(0, cam.height))
        L.orc_sample(slots.ctypes.data, len(slots), C.byref(cam), 1227, i, j, int(rng.integers(0, 1000)), rgb.ctypes.data, None)
    n = min(L.orc_logged_segments(), len(buf))
    L.orc_log_segments(None, 0)
    return buf[:n]


def grid_cells(seg, slots, cell=1.0):
    """Cells of a 2-D (x, z) grid a segment crosses inside the slab of the small spheres, up to its hit distance."""
    r = np.abs(slots["r"])
    small = (r > 0) & (r < 5 * np.median(r))
    c, rs = slots["c"][small].astype(np.float64), r[small].astype(np.float64)
    lo = (c - rs[:, None]).min(axis=0)
    hi = (c + rs[:, None]).max(axis=0)
    o, d, t_hit = seg[:, 0:3].astype(np.float64), seg[:, 3:6].astype(np.float64), seg[:, 6].astype(np.float64)
    t0 = np.zeros(len(seg))
    t1 = np.where(np.isfinite(t_hit), t_hit, 1e30)
    with np.errstate(divide="ignore", invalid="ignore"):
        for ax in range(3):                                   # clip to the box [lo, hi] of the small spheres
            inv = 1.0 / d[:, ax]
            ta, tb = (lo[ax] - o[:, ax]) * inv, (hi[ax] - o[:, ax]) * inv
            near, far = np.minimum(ta, tb), np.maximum(ta, tb)
            par = d[:, ax] == 0
            inside = (o[:, ax] >= lo[ax]) & (o[:, ax] <= hi[ax])
            near = np.where(par, np.where(inside, -np.inf, np.inf), near)
            far = np.where(par, np.where(inside, np.inf, -np.inf), far)
            t0, t1 = np.maximum(t0, near), np.minimum(t1, far)
    ok = t1 >= t0
    p0, p1 = o + d * t0[:, None], o + d * np.where(ok, t1, t0)[:, None]
    cells = 1 + np.abs(np.floor(p1[:, 0] / cell) - np.floor(p0[:, 0] / cell)) + np.abs(np.floor(p1[:, 2] / cell) - np.floor(p0[:, 2] / cell))
    return np.where(ok, cells, 0), int(small.sum())


def report(name, slots, cam, n_paths):
    rng = np.random.default_rng(1)
    seg = segments_of(slots, cam, n_paths, rng)
    sec = seg[seg[:, 8] > 0]                                  # scattered segments (camera rays go through the tile lists)
    cells, n_small = grid_cells(sec, slots)
    groups = cells[: len(cells) // 32 * 32].reshape(-1, 32)
    q = np.percentile(cells, [50, 90, 99])
    print(f"{name}: {len(slots)} slots ({n_small} in the grid), {n_paths} paths, {len(seg) / n_paths:.2f} segments/path, "
          f"{len(sec)} scattered segments")
    print(f"  cells per scattered segment: mean {cells.mean():.2f}, median {q[0]:.0f}, p90 {q[1]:.0f}, p99 {q[2]:.0f}, max {cells.max():.0f}; "
          f"{(cells == 0).mean() * 100:.0f} % never enter the slab")
    print(f"  per group of 32 segments (lock step): mean of the maximum {groups.max(axis=1).mean():.1f}, mean of the sum / 32 {groups.mean():.2f}")


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    report("scene 1", O.scene(1), O.camera(3840, 2160, 1000, 50), n)
    report("scaled scene (half 158)", O.scene_scaled(158), O.camera(3840, 2160, 256, 50), max(500, n // 20))
