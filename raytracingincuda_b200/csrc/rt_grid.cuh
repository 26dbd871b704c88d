// rt_grid.cuh -- uniform-grid closest hit for fields of similar spheres on a plane (the reference's scenes, BASELINE configs
// 2-5): what RT_ACCEL_AUTO picks for them since round 2.  The algorithm is stated operation by operation in float32 in
// tools/grid_model.py and checked on the CPU against the oracle's hit_world on logged path segments (tests/test_grid_model.py);
// on the GPU the kernel returns the linear scan's (slot id, t) bit for bit (tests/test_gpu_parity.py: primary passes and whole
// frames on scenes from 40 to 99 860 slots, random fields with odd-sized, negative-radius and duplicate spheres, float and double).
//
// A 2-D grid over the two long axes of the small spheres holds, per cell, the slots whose padded footprint overlaps the
// cell; a ray walks the cells of its projection (Amanatides-Woo) inside the inflated box of those spheres and runs the
// reference's exact sphere test (bvh_test_sphere) on what the cells list; spheres of a very different size (the ground, the
// three big ones, the zero-radius slot) are tested for every ray.  tools/analyse_accel.py: 1.3 cells per scattered segment on
// scene 1 and on the 99 860-slot scene alike, where the LBVH makes 6-13 node visits.
// Conservativeness (same argument as tools/grid_model.py): a sphere can only be hit if the ray passes within r + delta of its
// centre, delta = sqrt(rmin^2 + KEPS D^2) - rmin; footprints are registered with pad = 0.05 h; rays with delta <= pad / 2 walk
// the thin line, steps with a larger delta (cells hundreds of units from the origin) also look at k = ceil((delta - pad/2) / h)
// rings of cells around it (the hit point lies within r + delta of the centre per axis, so the registered footprint reaches
// into a cell at most that many cells from the hit point's); the walk stops after a cell whose exit parameter lies beyond
// the closest hit so far.
#pragma once
#include "rt_lbvh.cuh"

namespace rt {

#ifndef RT_GRID_RING_EDGE
#define RT_GRID_RING_EDGE 1
#endif
struct GridView {
    const unsigned int *start;    // [nu * nw + 1] first item of each cell
    const unsigned int *items;    // slots, cell by cell
    const float4 *big_geom;       // spheres outside the grid: tested for every ray
    const int *big_slot;
    int nbig;
    unsigned int n_items;         // entries of items[]
    int n_slots;                  // slots of the scene
    int nu, nw;
    int au, av, aw;               // axis numbers of the grid's u, of the slab, of the grid's w
    float lo[3], hi[3];           // bounds of the grid spheres (centre -/+ radius), rounded outwards, in x y z order
    float h, inv_h, pad, rmin;
    float ulo, wlo, half_pad;     // lo[au], lo[aw], pad / 2 (constant-bank operands instead of registers in the walk)
};

__constant__ GridView g_grid;     // one grid per device (set by the launch that uses it)

__device__ __forceinline__ float axis_of(const Vec3<float> &v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
__device__ __forceinline__ float4 ldg_geom(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ double4 ldg_geom(const double4 *p) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}
// the closest hit so far as a float that is not below it (the walk's stop test runs in float)
__device__ __forceinline__ float hit_t_up(float t) { return t; }
__device__ __forceinline__ float hit_t_up(double t) { return __double2float_ru(t); }

// Cold path of a step whose inflation reaches beyond its own cell (cells hundreds of units from the ray origin): every cell
// within k rings.  Out of line: the walk's hot loop stays small (the kernel is instruction-fetch bound).
// Consecutive ring steps overlap in all but their leading edge: after a step that tested the whole block around (cu', cw') with
// the same k, a move by one cell only brings the 2k+1 cells of the new column (or row) -- the others have been tested, and
// testing a sphere twice changes nothing (its first root > tmin is a function of the ray; the minimum has already seen it).
// `prev` remembers the last ring step (cell, k, step number); on the 99 860-slot scene the ring loops were 29 % of the
// kernel's instructions at 3 active threads (profiles/r02_pb_grid_100k_by_source_line.txt of build 7799f1692b92).
struct RingPrev { int cu, cw, k, step; };
template <typename T>
__device__ __noinline__ void grid_ring_tests(const typename Num<T>::vec4 *__restrict__ geom, int cu, int cw, int k, int step, RingPrev &prev,
                                             Vec3<T> o, Vec3<T> d, T a, Hit<T> &hit, unsigned &n_tests) {
    const GridView &g = g_grid;
    int b0 = max(cw - k, 0), b1 = min(cw + k, g.nw - 1), c0 = max(cu - k, 0), c1 = min(cu + k, g.nu - 1);
#if RT_GRID_RING_EDGE
    const bool chained = prev.step == step + 1 && prev.k == k;        // steps_left counts down: the previous step was a ring step too
    const int du = chained ? cu - prev.cu : 0, dw = chained ? cw - prev.cw : 0;
    prev.cu = cu; prev.cw = cw; prev.k = k; prev.step = step;
    if (chained) {
        if (du == 0 && dw == 0) return;                               // clamped at the border: the same block
        if (dw == 0 && (du == 1 || du == -1)) {
            c0 = c1 = cu + du * k;
            if (c0 < 0 || c0 >= g.nu) return;
        } else if (du == 0 && (dw == 1 || dw == -1)) {
            b0 = b1 = cw + dw * k;
            if (b0 < 0 || b0 >= g.nw) return;
        }
    }
#endif
    for (int b = b0; b <= b1; ++b)
        for (int c = c0; c <= c1; ++c) {
            const int cell = b * g.nu + c;
            RT_CHECK(cell >= 0 && cell < g.nu * g.nw, 601);
            const unsigned int e0 = __ldg(g.start + cell), e1 = __ldg(g.start + cell + 1);
            RT_CHECK(e0 <= e1 && e1 <= g.n_items, 602);
            for (unsigned int e = e0; e < e1; ++e) {
                const int slot = (int)__ldg(g.items + e);
                RT_CHECK(slot >= 0 && slot < g.n_slots, 603);
                bvh_test_sphere<T>(ldg_geom(geom + slot), slot, o, d, a, hit);
            }
            n_tests += e1 - e0;
        }
}

// closest hit of one ray; `geom` is the scene's geometry by slot (global memory).
// One loop tests "the current list" -- first the spheres outside the grid, then the cell of every step of the walk -- so the
// exact sphere test (sqrt and two IEEE divisions inline) exists once in the hot code.
// T = double (a GlobalDouble scene): the walk runs in float on the rounded ray -- rounding moves a point of the ray by ~1e-7
// relative, four orders below the registration padding, and the double discriminant's own noise is ~1e-16, so the cells the
// float walk visits list every sphere the double test can accept -- and the exact tests run in double on the double geometry.
template <typename T>
__device__ __forceinline__ Hit<T> grid_closest_hit(const GridView &g, const typename Num<T>::vec4 *__restrict__ geom, const Vec3<T> &oT,
                                                   const Vec3<T> &dT, unsigned &n_cells, unsigned &n_tests) {
    const float inf = Num<float>::inf();
    Hit<T> hit;
    hit.t = Num<T>::inf();
    hit.id = -1;
    const T aT = dot3(dT, dT);
    Vec3<float> o, d;
    o.x = (float)oT.x; o.y = (float)oT.y; o.z = (float)oT.z;
    d.x = (float)dT.x; d.y = (float)dT.y; d.z = (float)dT.z;
    const float a = sizeof(T) == 4 ? (float)aT : dot3(d, d);
    // walk state (set up after the first list)
    float t1 = 0.0f, iu_inv = 0.0f, iw_inv = 0.0f, delta = 0.0f;
    int iu = 0, iw = 0, su = 0, sw = 0, k_global = 0, steps_left = -1;        // steps_left < 0: the big list is being tested
    unsigned int e = 0u, e1 = (unsigned int)g.nbig;
    RingPrev ring_prev;                                                       // the other fields are read only after a ring step set them
    ring_prev.step = -2;                                                      // no ring step yet (steps_left + 1 is never -2)
    for (;;) {
        n_tests += e1 - e;
        const unsigned int *list = steps_left < 0 ? reinterpret_cast<const unsigned int *>(g.big_slot) : g.items;
#pragma unroll 1
        for (; e < e1; ++e) {
            const int slot = (int)__ldg(list + e);
            RT_CHECK(slot >= 0 && slot < g.n_slots, 603);
            bvh_test_sphere<T>(ldg_geom(geom + slot), slot, oT, dT, aT, hit);
        }
        if (steps_left < 0) {
            // ---- the spheres outside the grid are done: set up the walk ----
            // per-ray inflation, as bvh_start<true>
            const float fx = fmaxf(fabsf(g.lo[0] - o.x), fabsf(g.hi[0] - o.x));
            const float fy = fmaxf(fabsf(g.lo[1] - o.y), fabsf(g.hi[1] - o.y));
            const float fz = fmaxf(fabsf(g.lo[2] - o.z), fabsf(g.hi[2] - o.z));
            const float D2 = fz * fz + (fy * fy + fx * fx);
            const float fmx = fmaxf(fmaxf(fx, fy), fz);
            const float omax = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
            const float root = sqrt_approx(BVH_KEPS * D2 + g.rmin * g.rmin) * (1.0f + 2e-7f);
            delta = ((root - g.rmin) * 1.001f + 1e-7f) + 4.8e-7f * (omax + fmx);
            const float half_pad = g.half_pad;
            k_global = delta <= half_pad ? 0 : (int)fminf(ceilf((delta - half_pad) / g.h), 8192.0f);    // a NaN ray leaves at the clip test below
            const float infl = delta + 1e-6f * (omax + fmx);
            // clip to the inflated box of the grid spheres; |1/d| <= 1e30 keeps every product finite (rcp.rn == 1.0f / x)
            Vec3<float> inv;
            inv.x = d.x != 0.0f ? fminf(fmaxf(__frcp_rn(d.x), -1e30f), 1e30f) : 1e30f;
            inv.y = d.y != 0.0f ? fminf(fmaxf(__frcp_rn(d.y), -1e30f), 1e30f) : 1e30f;
            inv.z = d.z != 0.0f ? fminf(fmaxf(__frcp_rn(d.z), -1e30f), 1e30f) : 1e30f;
            float t0 = 0.0f;
            t1 = hit_t_up(hit.t);
            {
                const float ax = ((g.lo[0] - infl) - o.x) * inv.x, bx = ((g.hi[0] + infl) - o.x) * inv.x;
                const float ay = ((g.lo[1] - infl) - o.y) * inv.y, by = ((g.hi[1] + infl) - o.y) * inv.y;
                const float az = ((g.lo[2] - infl) - o.z) * inv.z, bz = ((g.hi[2] + infl) - o.z) * inv.z;
                t0 = fmaxf(fmaxf(t0, fminf(ax, bx)), fmaxf(fminf(ay, by), fminf(az, bz)));
                t1 = fminf(fminf(t1, fmaxf(ax, bx)), fminf(fmaxf(ay, by), fmaxf(az, bz)));
            }
            if (!(t0 <= t1 * 1.0001f + 1e-6f)) break;
            t0 = fmaxf(0.0f, t0 - 1e-4f * fabsf(t0) - 1e-6f);
            const float ou = axis_of(o, g.au), ow = axis_of(o, g.aw);
            const float du = axis_of(d, g.au), dw = axis_of(d, g.aw);
            iu_inv = axis_of(inv, g.au); iw_inv = axis_of(inv, g.aw);
            const float ulo = g.ulo, wlo = g.wlo;
            const float pu = ou + du * t0, pw = ow + dw * t0;
            // walking indices are not clamped (the inflated clip box reaches a little beyond the grid); look-ups are
            iu = (int)fminf(fmaxf(floorf((pu - ulo) * g.inv_h), -65536.0f), 65536.0f);
            iw = (int)fminf(fmaxf(floorf((pw - wlo) * g.inv_h), -65536.0f), 65536.0f);
            su = du > 0.0f ? 1 : (du < 0.0f ? -1 : 0);
            sw = dw > 0.0f ? 1 : (dw < 0.0f ? -1 : 0);
            steps_left = 2 * (g.nu + g.nw) + 64;
        } else {
            // ---- the cell of this step is done: stop, or move on to the next cell of the line ----
            // exit parameters of the current cell, recomputed from the cell index (no accumulated drift)
            const float tu = su == 0 ? inf : ((g.ulo + (float)(iu + (su > 0 ? 1 : 0)) * g.h) - axis_of(o, g.au)) * iu_inv;
            const float tw = sw == 0 ? inf : ((g.wlo + (float)(iw + (sw > 0 ? 1 : 0)) * g.h) - axis_of(o, g.aw)) * iw_inv;
            const float t_exit = fminf(tu, tw), stop = fminf(hit_t_up(hit.t), t1);
            if (!(t_exit <= stop * 1.0001f + 1e-6f) || --steps_left <= 0) break;
            if (tu <= tw) iu += su; else iw += sw;
        }
        // ---- the list of the step's cell ----
        ++n_cells;
        const int cu = min(max(iu, 0), g.nu - 1), cw = min(max(iw, 0), g.nw - 1);
        int k = 0;
        if (k_global > 0) {
            // Inflation of THIS step: a sphere whose root lies in the current cell is at most t_far |d| + 2 h + delta away from
            // the origin (t_far: where the ray leaves the cell or the walk ends).  With one inflation per ray (evaluated at the
            // far corner of the grid) every step of the 99 860-slot scene looked at rings of cells.  Rays whose inflation at the
            // far corner already stays inside the padding (k_global == 0: every ray of a compact scene) skip this.
            const float tu = su == 0 ? inf : ((g.ulo + (float)(iu + (su > 0 ? 1 : 0)) * g.h) - axis_of(o, g.au)) * iu_inv;
            const float tw = sw == 0 ? inf : ((g.wlo + (float)(iw + (sw > 0 ? 1 : 0)) * g.h) - axis_of(o, g.aw)) * iw_inv;
            const float t_far = fminf(fminf(tu, tw), t1);
            const float length = sqrtf(a), omax = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
            const float Ds = (t_far * length) * 1.0001f + (2.0f * g.h + delta);
            const float rs = sqrt_approx(BVH_KEPS * (Ds * Ds) + g.rmin * g.rmin) * (1.0f + 2e-7f);
            const float ds = ((rs - g.rmin) * 1.001f + 1e-7f) + 4.8e-7f * (omax + Ds);
            k = ds <= g.half_pad ? 0 : min((int)fminf(ceilf((ds - g.half_pad) / g.h), 8192.0f), k_global);
        }
        if (k > 0) {
            grid_ring_tests<T>(geom, cu, cw, k, steps_left, ring_prev, oT, dT, aT, hit, n_tests);
            e = e1 = 0u;
        } else {
            const int cell = cw * g.nu + cu;
            RT_CHECK(cell >= 0 && cell < g.nu * g.nw, 601);
            e = __ldg(g.start + cell);
            e1 = __ldg(g.start + cell + 1);
            RT_CHECK(e <= e1 && e1 <= g.n_items, 602);
        }
    }
    return hit;
}

}  // namespace rt
