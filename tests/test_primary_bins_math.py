"""CPU check of the camera-ray bin criterion (raytracingincuda_b200/csrc/rt_primary_bins.cuh: pb_bundle / pb_touches).

The kernels only run on a GPU; the MATH they rest on can be checked here: a numpy restatement of the criterion builds the
candidate list of a tile, camera rays of that tile are generated at the extremes of their two random draws (pixel jitter
at the corners of the pixel, lens sample on the rim of the defocus disk) plus random ones, rounded to float like
camera_ray does, and the oracle's hit_world (the reference's exact float scan, GF hittable.h:80-98) says which slot each
ray hits.  That slot must be in the tile's list -- for every ray, on the reference's scenes, on shifted / scaled copies
and on random sphere soups, at the tile geometry of a 320x192 and of a 3840x2160 frame.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

PB_SHIFT, PB_CAP = 4, 63
PB_NOISE = 64.0 * 2.0 ** -24


def bundle(cam, tx, ty):
    """pb_bundle: axis, opening and margins of the camera rays of tile (tx, ty), in double."""
    W, H = cam.width, cam.height
    i0, j0 = tx << PB_SHIFT, ty << PB_SHIFT
    i1, j1 = min(i0 + (1 << PB_SHIFT), W) - 1, min(j0 + (1 << PB_SHIFT), H) - 1
    mx, my = 0.5 * (i0 + i1), 0.5 * (j0 + j1)
    hx, hy = 0.5 * (i1 - i0) + 0.5 + 1e-3, 0.5 * (j1 - j0) + 0.5 + 1e-3
    L0 = np.array(cam.center, dtype=np.float64)
    du, dv, p0 = (np.array(v, dtype=np.float64) for v in (cam.du, cam.dv, cam.pixel00))
    Q0 = p0 + mx * du + my * dv
    a = Q0 - L0
    la = np.linalg.norm(a)
    rho = 0.0
    if not cam.defocus_angle <= 0:
        U, V = np.array(cam.disk_u, dtype=np.float64), np.array(cam.disk_v, dtype=np.float64)
        uu, vv, uv = U @ U, V @ V, U @ V
        rho = np.sqrt(0.5 * (uu + vv + np.sqrt((uu - vv) ** 2 + 4 * uv * uv))) * (1 + 1e-9)
    margin = 1e-5 * (1 + np.linalg.norm(L0) + np.linalg.norm(Q0))
    hT = hx * np.linalg.norm(du) + hy * np.linalg.norm(dv)
    kappa = (hT + rho + margin) / la * (1 + 1e-6)
    assert la > 0 and kappa < 0.5
    return dict(L0=L0, ah=a / la, kappa=kappa, cosk=np.sqrt(1 - kappa * kappa), rho=rho, margin=margin, la=la, hT=hT,
                box=(i0, i1, j0, j1))


def tile_list(B, slots):
    """bin_kernel: the slots pb_touches accepts (sphere case: radius = rmin = r)."""
    c = slots["c"].astype(np.float64)
    r = np.abs(slots["r"].astype(np.float64))
    b = c - B["L0"]
    bb = (b * b).sum(axis=1)
    along = b @ B["ah"]
    perp = np.sqrt(np.maximum(bb - along * along, 0.0))
    D = np.sqrt(bb) + r + B["rho"] + B["margin"]
    reach = r + (np.sqrt(r * r + PB_NOISE * D * D) - r)
    R0 = reach + B["rho"] + B["margin"]
    lhs, rhs = perp * B["cosk"], (R0 + B["kappa"] * along) * (1 + 1e-9) + B["margin"]
    first = ~(lhs > rhs)
    u_lo = np.maximum((along - R0) / (1 + B["kappa"]), 0.0)
    u_hi = np.maximum((along + R0) / (1 - B["kappa"]), 0.0)
    rl, hl = (B["rho"] + B["margin"]) * (1 + 1e-6), (B["hT"] + B["margin"]) * (1 + 1e-6)
    s_lo, s_hi = u_lo / B["la"], u_hi / B["la"]
    w_max = np.maximum(np.abs(1 - s_lo) * rl + s_lo * hl, np.abs(1 - s_hi) * rl + s_hi * hl)
    second = ~(perp > (reach + w_max) * (1 + 1e-9) + B["margin"])
    return set(np.nonzero(first & second)[0].tolist())


def touches(B, c, radius, rmin):
    """pb_touches for one ball (a sphere: radius = rmin = r; the bounding ball of a BVH box: half diagonal, smallest radius inside)."""
    b = np.asarray(c, dtype=np.float64) - B["L0"]
    bb = b @ b
    along = b @ B["ah"]
    perp = np.sqrt(max(bb - along * along, 0.0))
    D = np.sqrt(bb) + radius + B["rho"] + B["margin"]
    reach = radius + (np.sqrt(rmin * rmin + PB_NOISE * D * D) - rmin)
    R0 = reach + B["rho"] + B["margin"]
    if perp * B["cosk"] > (R0 + B["kappa"] * along) * (1 + 1e-9) + B["margin"]:
        return False
    u_lo, u_hi = max((along - R0) / (1 + B["kappa"]), 0.0), max((along + R0) / (1 - B["kappa"]), 0.0)
    rl, hl = (B["rho"] + B["margin"]) * (1 + 1e-6), (B["hT"] + B["margin"]) * (1 + 1e-6)
    s_lo, s_hi = u_lo / B["la"], u_hi / B["la"]
    w_max = max(abs(1 - s_lo) * rl + s_lo * hl, abs(1 - s_hi) * rl + s_hi * hl)
    return not perp > (reach + w_max) * (1 + 1e-9) + B["margin"]


def tile_rays(cam, B, rng, n_random):
    """Camera rays of the tile the way camera_ray builds them (float fma chains), at the extremes of the two draws and at
    random ones.  Yields (o, d) as float32 triples."""
    f = np.float32
    i0, i1, j0, j1 = B["box"]
    p00, du, dv = (np.array(v, dtype=f) for v in (cam.pixel00, cam.du, cam.dv))
    ctr, U, V = (np.array(v, dtype=f) for v in (cam.center, cam.disk_u, cam.disk_v))

    def fma(a, b, c):            # float32 fused multiply-add (the double product of two floats is exact)
        return f(np.float64(a) * np.float64(b) + np.float64(c))

    corners = [(i, j, ox, oy) for i in (i0, i1) for j in (j0, j1) for ox in (-0.5, 0.5) for oy in (-0.5, 0.5)]
    rim = [(np.cos(t), np.sin(t)) for t in np.linspace(0, 2 * np.pi, 8, endpoint=False)] + [(0.0, 0.0)]
    cases = [(c, q) for c in corners for q in rim]
    for _ in range(n_random):
        i, j = int(rng.integers(i0, i1 + 1)), int(rng.integers(j0, j1 + 1))
        ang, rad = rng.uniform(0, 2 * np.pi), np.sqrt(rng.uniform(0, 1))
        cases.append(((i, j, rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5)), (rad * np.cos(ang), rad * np.sin(ang))))
    for (i, j, ox, oy), (q0, q1) in cases:
        px, py = f(f(i) + f(ox)), f(f(j) + f(oy))
        q0, q1 = f(q0 * 0.99999), f(q1 * 0.99999)            # q0^2 + q1^2 < 1 (GF vec3.h:112)
        target = np.array([fma(py, dv[k], fma(px, du[k], p00[k])) for k in range(3)], dtype=f)
        o = ctr.copy()
        if not cam.defocus_angle <= 0:
            o = np.array([fma(q1, V[k], fma(q0, U[k], ctr[k])) for k in range(3)], dtype=f)
        yield o, (target - o).astype(f)


def hit_slot(slots, o, d):
    t = C.c_float(0)
    return O.lib().orc_hit_world(slots.ctypes.data, len(slots), (C.c_float * 3)(*o), (C.c_float * 3)(*d), C.c_float(0.001),
                                 C.c_float(np.inf), C.byref(t))


def moved(slots, scale, shift):
    s = slots.copy()
    s["c"] = (s["c"].astype(np.float64) * scale + np.asarray(shift, dtype=np.float64)).astype(np.float32)
    s["r"] = (s["r"].astype(np.float64) * scale).astype(np.float32)
    return s


def soup(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(60, 400))
    s = np.zeros(n, dtype=O.SLOT_DTYPE)
    s["c"] = rng.uniform(-12, 14, size=(n, 3)).astype(np.float32)
    s["r"] = (10 ** rng.uniform(-1.5, 0.5, size=n)).astype(np.float32)
    return s


SCENES = {
    "scene1": lambda: O.scene(1), "scene2": lambda: O.scene(2), "scene3": lambda: O.scene(3),
    "shifted": lambda: moved(O.scene(1), 1.0, (37.0, 3.0, -21.0)),
    "half_size": lambda: moved(O.scene(1), 0.5, (0.0, 0.0, 0.0)),
    "soup1": lambda: soup(1), "soup2": lambda: soup(2),
}


@pytest.mark.parametrize("name", sorted(SCENES))
@pytest.mark.parametrize("frame", [(320, 192), (3840, 2160)])
def test_every_camera_ray_hits_a_slot_of_its_tile_list(name, frame):
    slots = SCENES[name]()
    cam = O.camera(*frame)
    rng = np.random.default_rng(hash((name, frame)) % (1 << 32))
    tiles_x, tiles_y = (cam.width + 15) >> 4, (cam.height + 15) >> 4
    n_rays = n_hits = 0
    sizes = []
    for _ in range(40):
        tx, ty = int(rng.integers(0, tiles_x)), int(rng.integers(0, tiles_y))
        B = bundle(cam, tx, ty)
        allowed = tile_list(B, slots)
        sizes.append(len(allowed))
        for o, d in tile_rays(cam, B, rng, 24):
            s = hit_slot(slots, o, d)
            n_rays += 1
            if s >= 0:
                n_hits += 1
                assert s in allowed, (name, frame, tx, ty, s, sorted(allowed))
    assert n_rays > 6000 and n_hits > 0.2 * n_rays
    # the criterion must also stay selective on the reference's scenes: short lists, (almost) no overflow
    if name.startswith("scene"):
        assert np.mean(sizes) < 12 and np.mean(np.array(sizes) > PB_CAP) < 0.1, sizes


def test_list_is_tight_around_a_single_sphere():
    """One sphere on the optical axis: tiles whose bundle passes it by more than a few radii must not list it."""
    cam = O.camera(3840, 2160)
    s = np.zeros(1, dtype=O.SLOT_DTYPE)
    s["c"][0] = (0.0, 0.0, 0.0)                      # the look-at point, in focus
    s["r"][0] = 0.2
    ids, _ = O.primary(s, O.camera(240, 135))       # one primary ray per 16 x 16 tile of the 4K frame (centre of its first pixel)
    listed = np.zeros((135, 240), dtype=bool)
    for ty in range(135):
        for tx in range(240):
            listed[ty, tx] = 0 in tile_list(bundle(cam, tx, ty), s)
    assert listed[ids >= 0].all()                    # conservative
    assert listed.sum() <= 2.5 * max(1, (ids >= 0).sum()) + 40      # and not much more than the silhouette plus a rim of tiles


@pytest.mark.parametrize("frame", [(320, 192), (3840, 2160)])
def test_a_box_is_visited_when_a_sphere_inside_it_is_hit(frame):
    """bin_kernel_bvh prunes a subtree when the bounding ball of its box (float corners, as the LBVH build rounds them) fails
    pb_touches with the subtree's smallest radius.  Whenever a camera ray of the tile hits a sphere, every box that contains
    that sphere must pass -- checked on boxes around the k nearest neighbours of the hit sphere, k = 1 .. 64."""
    slots = O.scene(1)
    cam = O.camera(*frame)
    rng = np.random.default_rng(frame[0])
    c, r = slots["c"].astype(np.float32), np.abs(slots["r"].astype(np.float32))
    small = np.nonzero((r > 0.05) & (r < 5))[0]                          # the tree's spheres (ground and zero-radius slot stay outside)
    tiles_x, tiles_y = (cam.width + 15) >> 4, (cam.height + 15) >> 4
    checked = 0
    for _ in range(60):
        B = bundle(cam, int(rng.integers(0, tiles_x)), int(rng.integers(0, tiles_y)))
        hit = {hit_slot(slots, o, d) for o, d in tile_rays(cam, B, rng, 8)}
        for s in hit & set(small.tolist()):
            order = small[np.argsort(((c[small] - c[s]) ** 2).sum(axis=1))]
            for k in (1, 2, 5, 16, 64):
                grp = order[:k]
                lo = (c[grp] - r[grp, None]).min(axis=0)                    # float32, like bvh_leaves_kernel / bvh_refit_kernel
                hi = (c[grp] + r[grp, None]).max(axis=0)
                ctr = 0.5 * (lo.astype(np.float64) + hi.astype(np.float64))
                ext = 0.5 * (hi.astype(np.float64) - lo.astype(np.float64))
                rb = np.sqrt(ext @ ext) * (1 + 1e-6) + 1e-6 * np.abs(ctr).sum()
                assert touches(B, ctr, rb, float(r[grp].min())), (frame, s, k)
                checked += 1
    assert checked > 200
