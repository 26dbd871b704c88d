// rt_lbvh.cuh -- on-GPU LBVH over the small spheres of a large scene (BASELINE config 5, the
// "optional LBVH built on-GPU for large sphere counts" of the north star) and a traversal whose
// result is the one the reference's linear hit_world loop (GF hittable.h:80-98) produces.
//
// Build (Karras 2012): Morton codes of the sphere centres -> cub radix sort -> one thread per
// internal node finds its range and split -> bottom-up refit with one atomic flag per node.
// Spheres that are huge relative to the scene (the ground sphere) are kept out of the tree in a
// short "big" list that every ray tests directly.
//
// Exactness.  The linear scan's answer is order independent: each sphere contributes
//   v(S) = first of its two roots that is > tmin,   answer = min v(S), ties -> lowest slot
// (the reference's shrinking-tmax test accepts S iff v(S) < closest).  The traversal evaluates
// v(S) with the reference's exact arithmetic (disc_of / roots) at the leaves, so it only has to
// be CONSERVATIVE about which leaves it visits.  In float the reference's discriminant carries an
// error of up to 2^-24 * (21 |oc|^2 + 7 r^2) * a, i.e. a sphere behaves as if its radius^2 were at most
// r^2 + 32*2^-24*D^2 for a ray whose origin is D away -- far-away rays see "noisy" hits around small
// spheres.  Child boxes are therefore inflated per ray by
//   delta = sqrt(rmin^2 + KEPS*D^2) - rmin        (KEPS = 32 * 2^-24)
// where D bounds the distance from the ray origin to the box and rmin is the smallest radius
// below the child; the slab comparison itself carries a 1e-5 relative slack for its own rounding
// and nodes are culled against best_t with a 1e-4 relative slack.
#pragma once
#include "rt_device.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace rt {

struct BvhView {
    const float4 *nodes;      // 4 x float4 per internal node (see node_store)
    const float4 *geom;       // [m] small spheres in Morton order
    const int *slot;          // [m] original slot of each sorted sphere
    const float4 *big_geom;   // [nbig] spheres kept out of the tree
    const int *big_slot;
    int m, nbig;
    // bounds of the tree's spheres (centre -/+ radius), its smallest radius, and whether the scene is compact enough for the
    // traversal with ONE inflation per ray (bvh_start<true>) instead of one square root per box
    float blo[3], bhi[3], rmin_all;
    int compact;
};

// 32 * 2^-24: the worst-case bound of the reference's discriminant error is 2^-24 * (21 |oc|^2 + 7 r^2) (DESIGN.md section 6,
// "Error bound" (i)); D bounds |oc| from above and, for a sphere inside the box, r as well
constexpr float BVH_KEPS = 32.0f * 5.9604645e-8f;
constexpr int BVH_STACK = 64;

// ------------------------------------------------------------------------------ build ------
__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void bvh_morton_kernel(const float4 *__restrict__ geom, const int *__restrict__ small_idx, int m,
                                  float3 lo, float3 inv_extent, uint32_t *__restrict__ keys, int *__restrict__ vals) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const int s = small_idx[k];
    const float4 g = geom[s];
    const float x = fminf(fmaxf((g.x - lo.x) * inv_extent.x * 1024.0f, 0.0f), 1023.0f);
    const float y = fminf(fmaxf((g.y - lo.y) * inv_extent.y * 1024.0f, 0.0f), 1023.0f);
    const float z = fminf(fmaxf((g.z - lo.z) * inv_extent.z * 1024.0f, 0.0f), 1023.0f);
    keys[k] = (expand_bits10((uint32_t)x) << 2) | (expand_bits10((uint32_t)y) << 1) | expand_bits10((uint32_t)z);
    vals[k] = s;
}

// boxes/rad are indexed by "node id": internal nodes 0..m-2, leaves m-1..2m-2
__global__ void bvh_leaves_kernel(const float4 *__restrict__ geom, const int *__restrict__ sorted_slot, int m,
                                  float4 *__restrict__ geom_sorted, float *__restrict__ box, float *__restrict__ rad) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const float4 g = geom[sorted_slot[k]];
    geom_sorted[k] = g;                               // the exact test keeps the signed radius (r*r, and 1/r in the normal)
    const float r = fabsf(g.w);                       // a negative radius is legal in the reference's arithmetic (hollow glass)
    float *b = box + 6 * (size_t)(m - 1 + k);
    b[0] = g.x - r; b[1] = g.y - r; b[2] = g.z - r;
    b[3] = g.x + r; b[4] = g.y + r; b[5] = g.z + r;
    rad[m - 1 + k] = r;
}

__device__ __forceinline__ int bvh_delta(const uint32_t *__restrict__ keys, int m, int i, int j) {
    if (j < 0 || j >= m) return -1;
    const uint32_t a = keys[i], b = keys[j];
    return a == b ? 32 + __clz((uint32_t)i ^ (uint32_t)j) : __clz(a ^ b);
}

// child encoding: >= 0 internal node, < 0 leaf (~value = sorted index)
__global__ void bvh_hierarchy_kernel(const uint32_t *__restrict__ keys, int m, int2 *__restrict__ children,
                                     int *__restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m - 1) return;
    const int d = bvh_delta(keys, m, i, i + 1) - bvh_delta(keys, m, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = bvh_delta(keys, m, i, i - d);
    int lmax = 2;
    while (bvh_delta(keys, m, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (bvh_delta(keys, m, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = bvh_delta(keys, m, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (bvh_delta(keys, m, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = (lo == gamma) ? ~gamma : gamma;
    const int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    parent[left < 0 ? (m - 1 + ~left) : left] = i;
    parent[right < 0 ? (m - 1 + ~right) : right] = i;
    if (i == 0) parent[0] = -1;
}

__global__ void bvh_refit_kernel(int m, const int2 *__restrict__ children, const int *__restrict__ parent,
                                 float *box, float *rad, int *flags, float4 *__restrict__ nodes) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    int node = parent[m - 1 + k];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&flags[node], 1) == 0) return;          // the sibling subtree is not done yet
        const int2 ch = children[node];
        const int li = ch.x < 0 ? (m - 1 + ~ch.x) : ch.x, ri = ch.y < 0 ? (m - 1 + ~ch.y) : ch.y;
        const volatile float *bl = box + 6 * (size_t)li, *br = box + 6 * (size_t)ri;
        float l[6], r[6];
        for (int q = 0; q < 6; ++q) { l[q] = bl[q]; r[q] = br[q]; }
        const float rl = ((volatile float *)rad)[li], rr = ((volatile float *)rad)[ri];
        float *bn = box + 6 * (size_t)node;
        for (int q = 0; q < 3; ++q) { bn[q] = fminf(l[q], r[q]); bn[q + 3] = fmaxf(l[q + 3], r[q + 3]); }
        rad[node] = fminf(rl, rr);
        // node record: lmin.xyz lmax.x | lmax.yz rmin.xy | rmin.z rmax.xyz | left right lrad rrad
        nodes[4 * (size_t)node + 0] = make_float4(l[0], l[1], l[2], l[3]);
        nodes[4 * (size_t)node + 1] = make_float4(l[4], l[5], r[0], r[1]);
        nodes[4 * (size_t)node + 2] = make_float4(r[2], r[3], r[4], r[5]);
        nodes[4 * (size_t)node + 3] = make_float4(__int_as_float(ch.x), __int_as_float(ch.y), rl, rr);
        node = parent[node];
    }
}

// ------------------------------------------------------------------------------ traversal ---
// exact sphere test with the order-independent acceptance rule (see the header comment); T = double for the uniform grid over
// a double scene (GD hittable.h:40-66)
template <typename T>
__device__ __forceinline__ void bvh_test_sphere(const typename Num<T>::vec4 s, int slot, const Vec3<T> &o, const Vec3<T> &d, T a, Hit<T> &hit) {
    using N = Num<T>;
    T h;
    const T disc = disc_of<T>(s, o, d, a, h);
    if (disc < T(0)) return;
    if (roots_below_tmin(h, disc, a)) return;
    const T sq = N::sqrt(disc);
    T v = N::div(N::sub(h, sq), a);
    if (!(N::tmin() < v)) {
        v = N::div(N::add(h, sq), a);
        if (!(N::tmin() < v)) return;
    }
    if (v < hit.t || (v == hit.t && slot < hit.id)) { hit.t = v; hit.id = slot; }
}
__device__ __forceinline__ void bvh_test_sphere(const float4 s, int slot, const Vec3<float> &o, const Vec3<float> &d, float a, Hit<float> &hit) {
    bvh_test_sphere<float>(s, slot, o, d, a, hit);
}

// entry parameter of the inflated box, or +inf when the ray cannot touch it before `limit`.
// The slabs are ((bound - o) -/+ delta) * inv -- subtract the origin first: fma(bound, inv, -o*inv) would cancel badly
// for origins far from the coordinate origin and break the conservative guarantee.
__device__ __forceinline__ float bvh_box_entry(float lx, float ly, float lz, float hx, float hy, float hz, float rmin,
                                               const Vec3<float> &o, const Vec3<float> &inv, float limit) {
    const float ax = lx - o.x, bx = hx - o.x, ay = ly - o.y, by = hy - o.y, az = lz - o.z, bz = hz - o.z;
    const float fx = fmaxf(fabsf(ax), fabsf(bx)), fy = fmaxf(fabsf(ay), fabsf(by)), fz = fmaxf(fabsf(az), fabsf(bz));
    const float D2 = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
    const float delta = fmaf(sqrt_approx(fmaf(BVH_KEPS, D2, rmin * rmin)) - rmin, 1.001f, 1e-7f);
    const float t0x = (ax - delta) * inv.x, t1x = (bx + delta) * inv.x;
    const float t0y = (ay - delta) * inv.y, t1y = (by + delta) * inv.y;
    const float t0z = (az - delta) * inv.z, t1z = (bz + delta) * inv.z;
    const float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
    const float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    const bool ok = tn <= fminf(fmaf(fabsf(tf), 1e-4f, tf) + 1e-30f, limit) && tf >= 0.0f;
    return ok ? tn : __int_as_float(0x7f800000);
}

// Same test with an inflation fixed per ray and the origin folded into the products: opi = (o + delta) * inv,
// omi = (o - delta) * inv, t = fma(bound, inv, -opi/omi) -- 6 FMAs per box, no square root (measured against 6 subtractions +
// 6 products: -2 % on scene 1).  The rounded product moves a slab plane by at most one ulp of |o|, which the 8 ulp of
// |o| + |box| inside the per-ray inflation (bvh_start<true>) already cover.  bvh_start<true> keeps |inv| <= 1e30 so that no
// product is infinite (see there).
__device__ __forceinline__ float bvh_box_entry_ray(float lx, float ly, float lz, float hx, float hy, float hz, const Vec3<float> &opi,
                                                   const Vec3<float> &omi, const Vec3<float> &inv, float limit) {
    const float t0x = fmaf(lx, inv.x, -opi.x), t1x = fmaf(hx, inv.x, -omi.x);
    const float t0y = fmaf(ly, inv.y, -opi.y), t1y = fmaf(hy, inv.y, -omi.y);
    const float t0z = fmaf(lz, inv.z, -opi.z), t1z = fmaf(hz, inv.z, -omi.z);
    const float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
    const float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    const bool ok = tn <= fminf(fmaf(fabsf(tf), 1e-4f, tf) + 1e-30f, limit) && tf >= 0.0f;
    return ok ? tn : __int_as_float(0x7f800000);
}

// Resumable traversal state of one ray.  The scalars live in registers; the stack (node id and entry
// parameter packed into one 64-bit word) lives in local memory (L1 resident) and survives across the
// turns of the persistent loop, so a warp can interleave "a few traversal steps for everybody" with
// shading/regeneration of the lanes that finished: the lanes of a warp need very different numbers of
// node visits (mean 18, tail > 100 for rays grazing the ground), and waiting for the slowest lane left
// 6 of 32 lanes active in the node loop.
struct BvhStack { unsigned long long e[BVH_STACK]; };
struct BvhTrav {
    int node;             // current internal node, -1 = no traversal in flight
    int sp;
    float a;
    Vec3<float> inv;      // 1/d
    Hit<float> hit;
    Vec3<float> op, om;   // compact scenes (bvh_start<true>): (o + delta) * inv, (o - delta) * inv with one inflation for every box
};

template <bool RAYD = false>
__device__ __forceinline__ void bvh_start(const BvhView &bv, const Vec3<float> &o, const Vec3<float> &d, BvhTrav &tv,
                                          unsigned &n_tests) {
    using N = Num<float>;
    tv.a = dot3(d, d);
    tv.hit.t = N::inf();
    tv.hit.id = -1;
    tv.node = -1;
    tv.sp = 0;
    // A ray whose |d|^2 is +inf or NaN hits nothing in the reference's arithmetic (GF hittable.h:40-66): both roots are
    // (h -/+ sqrt(disc)) / a with a = +inf or NaN, i.e. +-0 when the numerator is finite and NaN otherwise, and neither passes
    // `tmin < root`.  Such rays exist -- a path that lands on the zero-radius slot gets n = (p - c) * (1/0), and config 5 traces
    // 2 x 10^9 paths -- and with 1/d = +-0 every slab test below answers "inside": one lane then walked all 199 719 nodes and
    // tested all 99 860 leaves, ~100 ms at the end of a launch (profiles/logs/r02af_long_traversals.log).  Same answer, no walk.
    if (!(tv.a < N::inf())) return;
    // the spheres outside the tree -- and the tree itself when it is a single sphere -- through ONE copy of the exact test
    // (the traversal kernels are instruction-fetch bound: every inlined copy of the sqrt/div sequences costs I-cache)
    const int n_direct = bv.nbig + (bv.m == 1 ? 1 : 0);
#pragma unroll 1
    for (int b = 0; b < n_direct; ++b) {
        const bool big = b < bv.nbig;
        bvh_test_sphere(__ldg(big ? bv.big_geom + b : bv.geom), __ldg(big ? bv.big_slot + b : bv.slot), o, d, tv.a, tv.hit);
    }
    n_tests += n_direct;
    if (bv.m <= 1) return;
    tv.inv.x = __frcp_rn(d.x); tv.inv.y = __frcp_rn(d.y); tv.inv.z = __frcp_rn(d.z);      // == 1.0f / x, the shorter sequence
    tv.node = 0;
    if (RAYD) {
        // One inflation for the whole traversal: the per-box formula evaluated at the farthest corner of the TREE's bounds
        // and at the tree's smallest radius bounds every box's own inflation from above (it grows with the distance and
        // shrinks with the radius), so the traversal stays conservative for any origin; it is only tight -- and chosen by
        // the host -- when the scene is compact.  The extra 8 ulp of |o| + |box| cover the rounding of o +/- delta and of
        // the subtractions against it.
        const float fx = fmaxf(fabsf(bv.blo[0] - o.x), fabsf(bv.bhi[0] - o.x));
        const float fy = fmaxf(fabsf(bv.blo[1] - o.y), fabsf(bv.bhi[1] - o.y));
        const float fz = fmaxf(fabsf(bv.blo[2] - o.z), fabsf(bv.bhi[2] - o.z));
        const float D2 = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
        const float omax = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
        const float delta = fmaf(sqrt_approx(fmaf(BVH_KEPS, D2, bv.rmin_all * bv.rmin_all)) - bv.rmin_all, 1.001f, 1e-7f) +
                            4.8e-7f * (omax + fmaxf(fmaxf(fx, fy), fz));
        tv.op.x = o.x + delta; tv.op.y = o.y + delta; tv.op.z = o.z + delta;
        tv.om.x = o.x - delta; tv.om.y = o.y - delta; tv.om.z = o.z - delta;
        // The products below must stay finite: with an infinite 1/d (a zero direction component) fma(bound, inf, -inf) is NaN
        // for one plane of a slab and -inf for the other, and fminf/fmaxf would then REJECT a box the ray starts inside of
        // (seen twice in 96 million segments).  |1/d| <= 1e30 treats such a component as 1e-30: over any t that can reach the
        // scene the ray does not move along that axis either way, so inside/outside of the inflated slab decides, as it should.
        tv.inv.x = fminf(fmaxf(tv.inv.x, -1e30f), 1e30f);
        tv.inv.y = fminf(fmaxf(tv.inv.y, -1e30f), 1e30f);
        tv.inv.z = fminf(fmaxf(tv.inv.z, -1e30f), 1e30f);
        tv.op.x *= tv.inv.x; tv.op.y *= tv.inv.y; tv.op.z *= tv.inv.z;
        tv.om.x *= tv.inv.x; tv.om.y *= tv.inv.y; tv.om.z *= tv.inv.z;
    }
}

// one node visit; sets tv.node = -1 when the traversal is complete
template <bool RAYD = false>
__device__ __forceinline__ void bvh_step(const BvhView &bv, const Vec3<float> &o, const Vec3<float> &d, BvhTrav &tv, BvhStack &st,
                                         unsigned &n_nodes, unsigned &n_tests) {
    const float inf = Num<float>::inf();
    ++n_nodes;
    const int node = tv.node;
    RT_CHECK(node >= 0 && node < bv.m - 1 && tv.sp >= 0 && tv.sp <= BVH_STACK, 501);
    const float4 q0 = __ldg(bv.nodes + 4 * (size_t)node), q1 = __ldg(bv.nodes + 4 * (size_t)node + 1);
    const float4 q2 = __ldg(bv.nodes + 4 * (size_t)node + 2), q3 = __ldg(bv.nodes + 4 * (size_t)node + 3);
    const int left = __float_as_int(q3.x), right = __float_as_int(q3.y);
    const float limit = tv.hit.t * 1.0001f + 1e-6f;
    float tl, tr;
    if (RAYD) {
        tl = bvh_box_entry_ray(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tv.op, tv.om, tv.inv, limit);
        tr = bvh_box_entry_ray(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, tv.op, tv.om, tv.inv, limit);
    } else {
        tl = bvh_box_entry(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q3.z, o, tv.inv, limit);
        tr = bvh_box_entry(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.w, o, tv.inv, limit);
    }
    // leaf tests run right here (parking them per lane to run the tests of many lanes together was measured 4 % SLOWER: the
    // closest hit shrinks later, and the flush adds two ballots per step) -- both children through one copy of the test
    int leaf = -1, leaf2 = -1;
    if (tl < inf && left < 0) { leaf = ~left; tl = inf; }
    if (tr < inf && right < 0) { leaf2 = ~right; tr = inf; }
    if (leaf < 0) { leaf = leaf2; leaf2 = -1; }
#pragma unroll 1
    while (leaf >= 0) {
        RT_CHECK(leaf < bv.m, 503);
        bvh_test_sphere(__ldg(bv.geom + leaf), __ldg(bv.slot + leaf), o, d, tv.a, tv.hit);
        ++n_tests;
        leaf = leaf2;
        leaf2 = -1;
    }
    if (tl < inf && tr < inf) {
        const bool left_first = tl <= tr;
        // depth <= 30 Morton bits + the index tie-break of equal keys (log2 m <= 24): at most 54 entries; a full stack would
        // drop the far child silently, so stop loudly instead
        if (tv.sp >= BVH_STACK) __trap();
        st.e[tv.sp] = ((unsigned long long)__float_as_uint(left_first ? tr : tl) << 32) | (unsigned int)(left_first ? right : left);
        ++tv.sp;
        tv.node = left_first ? left : right;
        return;
    }
    if (tl < inf) { tv.node = left; return; }
    if (tr < inf) { tv.node = right; return; }
    // pop, skipping subtrees the current best hit already rules out
    tv.node = -1;
    const float keep = tv.hit.t * 1.0001f + 1e-6f;
    while (tv.sp > 0) {
        --tv.sp;
        const unsigned long long e = st.e[tv.sp];
        if (__uint_as_float((unsigned int)(e >> 32)) <= keep) { tv.node = (int)(unsigned int)e; break; }
    }
}

// whole traversal in one go (primary pass)
__device__ __forceinline__ Hit<float> bvh_closest_hit(const BvhView &bv, const Vec3<float> &o, const Vec3<float> &d,
                                                      unsigned &n_nodes, unsigned &n_tests) {
    BvhTrav tv;
    BvhStack st;
    bvh_start(bv, o, d, tv, n_tests);
    while (tv.node >= 0) bvh_step(bv, o, d, tv, st, n_nodes, n_tests);
    return tv.hit;
}

}  // namespace rt
