"""CPU check of the uniform-grid closest hit planned as RT_ACCEL_GRID (tools/grid_model.py).

Real path segments logged by the oracle (camera rays and scattered rays of scenes 1-3, of a shifted copy and of a scaled
scene) go through the float32 model of the grid walk; the closest hit it finds among the spheres it gets to test, plus the
spheres outside the grid, must be the oracle's full-scan hit -- same slot, same t, bit for bit -- while testing a few
spheres instead of hundreds.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle_lib as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import grid_model as GM  # noqa: E402


def logged_segments(slots, cam, n_paths, seed):
    L = O.lib()
    L.orc_log_segments.argtypes = [C.c_void_p, C.c_long]
    L.orc_logged_segments.restype = C.c_long
    buf = np.zeros((n_paths * 60, 9), dtype=np.float32)
    L.orc_log_segments(buf.ctypes.data, len(buf))
    rng = np.random.default_rng(seed)
    rgb = np.zeros(3, dtype=np.float32)
    for _ in range(n_paths):
        L.orc_sample(slots.ctypes.data, len(slots), C.byref(cam), 1227, int(rng.integers(0, cam.width)),
                     int(rng.integers(0, cam.height)), int(rng.integers(0, 1000)), rgb.ctypes.data, None)
    n = min(L.orc_logged_segments(), len(buf))
    L.orc_log_segments(None, 0)
    return buf[:n]


def closest_among(slots, idx, o, d):
    """The reference's exact closest hit over the slots `idx` (ascending): (t, slot) or (inf, -1)."""
    if len(idx) == 0:
        return np.float32(np.inf), -1
    idx = np.sort(np.asarray(idx, dtype=np.int64))
    sub = np.ascontiguousarray(slots[idx])
    t = C.c_float(0)
    k = O.lib().orc_hit_world(sub.ctypes.data, len(sub), (C.c_float * 3)(*o), (C.c_float * 3)(*d), C.c_float(0.001),
                              C.c_float(np.inf), C.byref(t))
    return (np.float32(t.value), int(idx[k])) if k >= 0 else (np.float32(np.inf), -1)


def moved(slots, scale, shift):
    s = slots.copy()
    s["c"] = (s["c"].astype(np.float64) * scale + np.asarray(shift, dtype=np.float64)).astype(np.float32)
    s["r"] = (s["r"].astype(np.float64) * scale).astype(np.float32)
    return s


SCENES = {
    "scene1": (lambda: O.scene(1), 400), "scene2": (lambda: O.scene(2), 300), "scene3": (lambda: O.scene(3), 300),
    "shifted": (lambda: moved(O.scene(1), 1.0, (37.0, 3.0, -21.0)), 300),
    "scaled24": (lambda: O.scene_scaled(24), 150),
    "scaled100": (lambda: O.scene_scaled(100), 60),       # 40 004 slots, 200 cells wide: far cells need rings, near ones do not
}


@pytest.mark.parametrize("name", sorted(SCENES))
def test_grid_walk_finds_the_full_scan_hit(name):
    make, n_paths = SCENES[name]
    slots = make()
    G = GM.Grid(slots)
    assert G.ok
    cam = O.camera(640, 360, 1000, 50)
    seg = logged_segments(slots, cam, n_paths, seed=len(slots))
    assert len(seg) > 2 * n_paths
    tested_total = cells_total = 0
    for row in seg:
        o, d = row[0:3], row[3:6]
        want_t, want_s = (np.float32(row[6]), int(row[7]))
        big_t, big_s = closest_among(slots, G.big, o, d)
        t, s, tested, cells = GM.candidates(G, o, d, big_t, lambda idx: closest_among(slots, idx, o, d))
        if big_t < t or (big_t == t and big_s >= 0 and (s < 0 or big_s < s)):
            t, s = big_t, big_s
        assert s == want_s, (name, o, d, s, want_s, sorted(tested))
        if want_s >= 0:
            assert np.float32(t).view(np.uint32) == want_t.view(np.uint32)
        tested_total += len(tested)
        cells_total += cells
    # the point of the structure: a handful of exact tests per segment, whatever the size of the scene
    assert tested_total / len(seg) < 6 and cells_total / len(seg) < 4, (tested_total / len(seg), cells_total / len(seg))


def test_one_inflation_per_ray_is_exact_too_but_wasteful_on_large_grids():
    """The variant that ran on hardware first (csrc/rt_grid.cuh in round 1): same hits, many more tests on a wide grid."""
    slots = O.scene_scaled(100)
    G = GM.Grid(slots)
    seg = logged_segments(slots, O.camera(640, 360, 1000, 50), 25, seed=3)
    tested = {True: 0, False: 0}
    for row in seg:
        o, d = row[0:3], row[3:6]
        big_t, big_s = closest_among(slots, G.big, o, d)
        res = {}
        for local in (True, False):
            t, s, used, _ = GM.candidates(G, o, d, big_t, lambda idx: closest_among(slots, idx, o, d), local=local)
            if big_t < t or (big_t == t and big_s >= 0 and (s < 0 or big_s < s)):
                t, s = big_t, big_s
            res[local] = (np.float32(t).view(np.uint32), s)
            tested[local] += len(used)
        assert res[True] == res[False] and res[True][1] == int(row[7])
    assert tested[False] > 2.5 * tested[True], tested


def test_far_origins_widen_the_walk():
    """Rays that start hundreds of units away see the small spheres through the float noise of the reference's
    discriminant (r_eff >> r): the model must still find what the full scan finds."""
    slots = O.scene(1)
    G = GM.Grid(slots)
    rng = np.random.default_rng(5)
    n_hit = 0
    for _ in range(400):
        ang, dist = rng.uniform(0, 2 * np.pi), rng.uniform(150, 900)
        o = np.array([dist * np.cos(ang), rng.uniform(0.05, 3.0), dist * np.sin(ang)], dtype=np.float32)
        target = np.array([rng.uniform(-11, 11), rng.uniform(0.0, 0.4), rng.uniform(-11, 11)], dtype=np.float32)
        d = ((target - o) * np.float32(rng.uniform(0.2, 2.0))).astype(np.float32)
        want_t, want_s = closest_among(slots, np.arange(len(slots)), o, d)
        big_t, big_s = closest_among(slots, G.big, o, d)
        t, s, tested, cells = GM.candidates(G, o, d, big_t, lambda idx: closest_among(slots, idx, o, d))
        if big_t < t or (big_t == t and big_s >= 0 and (s < 0 or big_s < s)):
            t, s = big_t, big_s
        assert s == want_s and (want_s < 0 or np.float32(t).view(np.uint32) == want_t.view(np.uint32)), (o, d, s, want_s)
        n_hit += want_s >= 0 and want_s not in set(G.big.tolist())
    assert n_hit > 50


def test_ring_steps_need_only_their_leading_edge():
    """grid_ring_tests (rt_grid.cuh, final round-2 build): after a ring step with the same k, a move by one cell brings the
    2k+1 cells of the new column or row -- the rest of the block was looked at by the previous step.  Same tested slots and
    the same hit as looking at the whole block every step, with a fraction of the look-ups: rays from far origins across
    scene 1 (several rings along the whole walk) and logged segments of a 40 004-slot field (rings for the far cells only)."""
    rng = np.random.default_rng(9)
    cases = []
    s1 = O.scene(1)
    G1 = GM.Grid(s1)
    for _ in range(150):
        ang, dist = rng.uniform(0, 2 * np.pi), rng.uniform(150, 900)
        o = np.array([dist * np.cos(ang), rng.uniform(0.05, 3.0), dist * np.sin(ang)], dtype=np.float32)
        target = np.array([rng.uniform(-11, 11), rng.uniform(0.0, 0.4), rng.uniform(-11, 11)], dtype=np.float32)
        cases.append((s1, G1, o, ((target - o) * np.float32(rng.uniform(0.2, 2.0))).astype(np.float32)))
    big = O.scene_scaled(100)
    Gb = GM.Grid(big)
    for row in logged_segments(big, O.camera(640, 360, 1000, 50), 40, seed=17):
        cases.append((big, Gb, row[0:3], row[3:6]))
    # camera rays towards the horizon of the large field: hundreds of steps, the far ones with rings
    for _ in range(40):
        o = np.array([13.0, 2.0, 3.0], dtype=np.float32)
        target = np.array([rng.uniform(-95, -40), rng.uniform(0.0, 0.4), rng.uniform(-95, 95)], dtype=np.float32)
        cases.append((big, Gb, o, (target - o).astype(np.float32)))
    look = {True: 0, False: 0}
    ring_rays = 0
    for slots, G, o, d in cases:
        big_t, _ = closest_among(slots, G.big, o, d)
        res = {}
        for edge in (True, False):
            st = {}
            t, s, tested, cells = GM.candidates(G, o, d, big_t, lambda idx: closest_among(slots, idx, o, d), ring_edge=edge, stats=st)
            res[edge] = (np.float32(t).view(np.uint32), s, frozenset(tested), cells)
            look[edge] += st.get("lookups", 0)
        assert res[True] == res[False], (o, d)
        ring_rays += res[True][3] > 0
    assert ring_rays > 100 and look[True] < 0.6 * look[False], look
