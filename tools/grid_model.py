"""Executable model of the uniform-grid closest hit planned as RT_ACCEL_GRID -- CPU only.

The reference's scenes are fields of equal small spheres on a plane.  A 2-D grid over the two long axes holds, per cell,
the spheres whose padded footprint overlaps it; a ray walks the cells of its projection (Amanatides-Woo) inside the slab
of the small spheres and runs the reference's exact sphere test on what the cells list; the few spheres of a very different
size (ground, the three big ones, the zero-radius slot) are tested for every ray.  The closest hit is order independent
(each sphere contributes its first root > tmin, the minimum wins, ties go to the lowest slot), so the walk only has to
visit a SUPERSET of the spheres whose reference discriminant is >= 0 and whose root lies before the point where it stops.

This file states the algorithm in float32 exactly as the kernel is meant to compute it (one rounding per operation, no
contraction), so that tests/test_grid_model.py can check it against the oracle's hit_world on real path segments before any
GPU time is spent.

Conservativeness.
  * Registration: sphere (c, r) goes into every cell that its footprint [c - (r + pad), c + (r + pad)]^2 overlaps,
    pad = PAD * h.
  * The reference's float discriminant can be >= 0 only if the ray passes within r + delta of c,
    delta = sqrt(rmin^2 + KEPS * D^2) - rmin (rt_lbvh.cuh), D = distance from the origin to the far corner of the grid.
    A ray with delta <= pad / 2 walks the thin line; the point where it enters the ball of radius r + delta lies pad / 2
    inside the registered footprint, far more than the rounding of the walk.  A ray with a larger delta (origin hundreds of
    units away) also looks at the k = ceil((delta - pad / 2) / h) rings of cells around every cell of the line: the hit
    point Q lies within r + delta of c per axis, so the registered footprint reaches to within delta - pad of Q, i.e. into a
    cell at most ceil((delta - pad) / h) cells from Q's; pad / 2 of the padding is kept as margin for the walk's rounding.
  * Termination: the walk stops after a cell whose exit parameter lies beyond the closest hit so far (or beyond the end of
    the clipped range); a sphere with a closer root has that root inside a cell that was already visited.
"""
import numpy as np

f32 = np.float32
KEPS = f32(32.0) * f32(5.9604645e-8)
PAD = 0.05
TMIN = f32(0.001)


class Grid:
    """Host-side build (double precision is fine here: it only decides which cells list which sphere)."""

    def __init__(self, slots):
        c = slots["c"].astype(np.float64)
        r = np.abs(slots["r"].astype(np.float64))
        n = len(slots)
        finite = np.isfinite(c).all(axis=1) & np.isfinite(r)
        r_med = np.median(r[finite & (r > 0)]) if (finite & (r > 0)).any() else 0.0
        in_grid = finite & (r >= 0.25 * r_med) & (r <= 4.0 * r_med) & (r > 0)
        self.big = np.nonzero(~in_grid)[0].astype(np.int32)
        g = np.nonzero(in_grid)[0].astype(np.int32)
        self.ok = len(g) >= 2 and len(self.big) <= 64
        if not self.ok:
            return
        ext = c[g].max(axis=0) - c[g].min(axis=0)
        self.v = int(np.argmin(ext))                                   # the thin axis: the slab
        self.u, self.w = [a for a in range(3) if a != self.v]
        lo3 = (c[g] - r[g, None]).min(axis=0)
        hi3 = (c[g] + r[g, None]).max(axis=0)
        area = max((hi3[self.u] - lo3[self.u]) * (hi3[self.w] - lo3[self.w]), 1e-30)
        h = np.sqrt(area / len(g))
        h = max(h, 2.0 * r[g].max() * 0.5)                             # a cell is at least one largest radius wide
        self.nu = int(min(max(np.ceil((hi3[self.u] - lo3[self.u]) / h), 1), 4096))
        self.nw = int(min(max(np.ceil((hi3[self.w] - lo3[self.w]) / h), 1), 4096))
        h = max((hi3[self.u] - lo3[self.u]) / self.nu, (hi3[self.w] - lo3[self.w]) / self.nw, h)
        self.h = f32(h * (1 + 1e-6))                                   # float cell size, rounded up: nu * h covers the bounds
        self.inv_h = f32(1.0) / self.h
        # float bounds of the grid spheres (centre -/+ radius), rounded outwards
        self.lo = np.nextafter(lo3.astype(np.float32), f32(-np.inf))
        self.hi = np.nextafter(hi3.astype(np.float32), f32(np.inf))
        self.rmin = f32(r[g].min())
        self.pad = f32(PAD * float(self.h))
        # registration (double, with a relative 1e-6 on the footprint)
        hh, ulo, wlo = float(self.h), float(self.lo[self.u]), float(self.lo[self.w])
        cells = [[] for _ in range(self.nu * self.nw)]
        for i in g:
            R = (r[i] + float(self.pad)) * (1 + 1e-6)
            u0 = int(np.clip(np.floor((c[i, self.u] - R - ulo) / hh), 0, self.nu - 1))
            u1 = int(np.clip(np.floor((c[i, self.u] + R - ulo) / hh), 0, self.nu - 1))
            w0 = int(np.clip(np.floor((c[i, self.w] - R - wlo) / hh), 0, self.nw - 1))
            w1 = int(np.clip(np.floor((c[i, self.w] + R - wlo) / hh), 0, self.nw - 1))
            for iw in range(w0, w1 + 1):
                for iu in range(u0, u1 + 1):
                    cells[iw * self.nu + iu].append(int(i))
        self.start = np.zeros(self.nu * self.nw + 1, dtype=np.int64)
        self.start[1:] = np.cumsum([len(x) for x in cells])
        self.items = np.array([s for x in cells for s in x], dtype=np.int32)

    def cell(self, iu, iw):
        k = iw * self.nu + iu
        return self.items[self.start[k]:self.start[k + 1]]


def candidates(G, o, d, limit, t_of, local=True, ring_edge=True, stats=None):
    """Full model: walk with early termination.  `t_of(slots)` returns the closest (t, slot) among `slots` with the
    reference's exact arithmetic (the oracle), or (inf, -1).  Returns (best_t, best_slot, tested slots, cells visited).

    local=False is the first transcription that ran on a B200 (csrc/rt_grid.cuh, round 1): ONE inflation per ray, evaluated
    at the far corner of the grid.  On the 99 860-slot scene that corner is 450 units away, delta = 0.45 and every step looks
    at two rings of cells (73 exact tests per segment, 225 ms where the LBVH takes 47).  local=True evaluates the inflation
    per step, at the distance the ray has reached: a sphere whose root lies in the current cell is at most
    t_far * |d| + 2 h + delta_global from the origin (t_far: where the ray leaves the cell or the walk ends).

    ring_edge=True is grid_ring_tests of the final round-2 build: a ring step that follows a ring step with the same k and has
    moved by one cell looks at the 2k+1 cells of its leading edge only (the rest of its block was looked at by the previous
    step).  `stats["lookups"]` counts the cells looked at; tests/test_grid_model.py checks that both settings test the same
    set of slots and find the same hit."""
    o = np.asarray(o, dtype=f32)
    d = np.asarray(d, dtype=f32)
    inf = f32(np.inf)
    # per-ray inflation, as bvh_start<true> computes it
    fx = [max(abs(f32(G.lo[a] - o[a])), abs(f32(G.hi[a] - o[a]))) for a in range(3)]
    D2 = f32(f32(fx[2] * fx[2]) + f32(f32(fx[1] * fx[1]) + f32(fx[0] * fx[0])))
    omax = max(abs(o[0]), abs(o[1]), abs(o[2]))
    root = f32(np.sqrt(f32(f32(KEPS * D2) + f32(G.rmin * G.rmin)))) * f32(1.0 + 2e-7)      # sqrt.approx is within 2 ulp
    delta = f32(f32(f32(root - G.rmin) * f32(1.001)) + f32(1e-7)) + f32(f32(4.8e-7) * f32(omax + max(fx)))
    half_pad = f32(G.pad * f32(0.5))
    k = 0 if delta <= half_pad else int(np.ceil(float(f32(delta - half_pad)) / float(G.h)))
    infl = f32(delta + f32(1e-6) * f32(omax + max(fx)))
    # clip the ray to the inflated box of the grid spheres
    t0, t1 = f32(0), f32(limit)
    inv = [f32(0)] * 3
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for a in range(3):
            inv[a] = f32(min(max(f32(1.0) / d[a], f32(-1e30)), f32(1e30))) if d[a] != 0 else f32(1e30)
            ta = f32(f32(f32(G.lo[a] - infl) - o[a]) * inv[a])
            tb = f32(f32(f32(G.hi[a] + infl) - o[a]) * inv[a])
            near, far = min(ta, tb), max(ta, tb)
            t0, t1 = max(t0, near), min(t1, far)
    tested, cells = set(), 0
    best_t, best_s = inf, -1
    if not (t0 <= f32(t1 * f32(1.0001)) + f32(1e-6)):
        return best_t, best_s, tested, cells
    t0 = max(f32(0), f32(t0 - f32(f32(1e-4) * abs(t0)) - f32(1e-6)))
    pu = f32(o[G.u] + f32(d[G.u] * t0))
    pw = f32(o[G.w] + f32(d[G.w] * t0))
    # walking indices are not clamped (the inflated clip box reaches a little beyond the grid); look-ups are
    iu = int(np.clip(np.floor(float(f32(f32(pu - G.lo[G.u]) * G.inv_h))), -65536, 65536))
    iw = int(np.clip(np.floor(float(f32(f32(pw - G.lo[G.w]) * G.inv_h))), -65536, 65536))
    su = 1 if d[G.u] > 0 else (-1 if d[G.u] < 0 else 0)
    sw = 1 if d[G.w] > 0 else (-1 if d[G.w] < 0 else 0)

    def exit_t(i, s, axis):
        if s == 0:
            return inf
        edge = f32(G.lo[axis] + f32(f32(i + (1 if s > 0 else 0)) * G.h))
        return f32(f32(edge - o[axis]) * inv[axis])

    k_global = k
    length = f32(np.sqrt(f32(f32(d[2] * d[2]) + f32(f32(d[0] * d[0]) + f32(d[1] * d[1])))))
    prev = None                                                        # last ring step: (cu, cw, k, step)
    for step in range(2 * (G.nu + G.nw) + 64):
        cells += 1
        if local:
            t_far = min(min(exit_t(iu, su, G.u), exit_t(iw, sw, G.w)), t1)
            Ds = f32(f32(f32(t_far * length) * f32(1.0001)) + f32(f32(f32(2.0) * G.h) + delta))
            rs = f32(np.sqrt(f32(f32(KEPS * f32(Ds * Ds)) + f32(G.rmin * G.rmin)))) * f32(1.0 + 2e-7)
            ds = f32(f32(f32(rs - G.rmin) * f32(1.001)) + f32(1e-7)) + f32(f32(4.8e-7) * f32(omax + Ds))
            k = 0 if ds <= half_pad else min(int(np.ceil(float(f32(ds - half_pad)) / float(G.h))), k_global)
        cu, cw = min(max(iu, 0), G.nu - 1), min(max(iw, 0), G.nw - 1)
        b0, b1, c0, c1 = max(cw - k, 0), min(cw + k, G.nw - 1), max(cu - k, 0), min(cu + k, G.nu - 1)
        if ring_edge and k > 0:
            chained = prev is not None and prev[3] == step - 1 and prev[2] == k
            du, dw = (cu - prev[0], cw - prev[1]) if chained else (0, 0)
            prev = (cu, cw, k, step)
            if chained:
                if du == 0 and dw == 0:
                    b1 = b0 - 1                                         # clamped at the border: the same block again
                elif dw == 0 and abs(du) == 1:
                    c0 = c1 = cu + du * k
                    if not 0 <= c0 < G.nu:
                        c1 = c0 - 1
                elif du == 0 and abs(dw) == 1:
                    b0 = b1 = cw + dw * k
                    if not 0 <= b0 < G.nw:
                        b1 = b0 - 1
        for b in range(b0, b1 + 1):
            for a in range(c0, c1 + 1):
                if stats is not None:
                    stats["lookups"] = stats.get("lookups", 0) + 1
                new = [int(s) for s in G.cell(a, b) if int(s) not in tested]
                if new:
                    tested.update(new)
                    t, s = t_of(new)
                    if t < best_t or (t == best_t and s < best_s):
                        best_t, best_s = t, s
        tu, tw = exit_t(iu, su, G.u), exit_t(iw, sw, G.w)
        t_exit = min(tu, tw)
        stop = min(best_t, t1)
        if not (t_exit <= f32(stop * f32(1.0001)) + f32(1e-6)):
            break
        if tu <= tw:
            iu += su
        else:
            iw += sw
    return best_t, best_s, tested, cells
