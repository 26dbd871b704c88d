// rt_kernels.cu -- the sm_100a kernels of the render path and the device half of the C ABI.
//
//   trace_kernel     persistent-thread path tracer.  One lane = one path at a time; a lane whose
//                    path ends regenerates the next sample of its (pixel, chunk) job in place, a
//                    lane whose job ends pulls the next job from a global queue with one
//                    warp-aggregated atomic (ballot + popc), so all 32 lanes stay inside the
//                    sphere scan.  Replaces init_rng + render (GF rtweekend.h:43-50,
//                    GF camera.h:78-172).
//   trace_kernel_pb  (rt_primary_bins.cuh, the default) the same persistent tracer with the camera rays resolved against
//                    per-tile candidate lists that bin_kernel / bin_kernel_bvh build on the device before the frame;
//                    scattered rays go through the shared-memory scan or the LBVH.  Same image, bit for bit.
//   finalize_kernel  adds the fixed-point accumulators (of one or several GPUs), scales, gamma-encodes and writes
//                    the frame with 16-byte vector stores (GF camera.h:167-171, color.h:10-13).
//   primary_kernel   deterministic primary-ray (slot id, t) pass through the same closest-hit
//                    routine (parity probe for GF hittable.h:80-98).
#include "rt_b200.h"
#include "rt_device.cuh"
#include "rt_lbvh.cuh"
#include "rt_grid.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <vector>

namespace rt {

constexpr unsigned FULL = 0xffffffffu;
constexpr int ACCEL_LBVH_COMPACT = 3;      // internal: RT_ACCEL_LBVH with one box inflation per ray (compact scenes, rt_lbvh.cuh)
#ifndef RT_TRACE_BLOCK
#define RT_TRACE_BLOCK 256
#endif
constexpr int TRACE_BLOCK = RT_TRACE_BLOCK;     // threads per CTA of the path tracer (a multiple of 32)

#ifndef RT_TRACE_MIN_BLOCKS
#define RT_TRACE_MIN_BLOCKS 3
#endif

template <typename T> struct DevCamera {
    Vec3<T> center, pixel00, du, dv, disk_u, disk_v;
    T defocus_angle;
    T scale;
};

// Job = (pixel, sample range).  The image does not depend on how the samples are cut into jobs (integer accumulation,
// rt_device.cuh), so the cut is a pure scheduling choice (JobPlan, plan_jobs below): pixels are handed out in BANDS of rows
// small enough that a band's accumulators stay in L2 while its jobs run, bottom band first (the jobs that run while the machine
// drains are sky pixels); inside a band chunk-major: all pixels for the first sample range, then for the second, ...
// A job is ONE sample (spj = 1) unless spp is beyond 65 536: the lanes of a warp that finish a path in the same turn claim
// consecutive jobs -- adjacent pixels, same sample index -- so their camera rays, their hits and their scattered rays stay
// close to each other (same tile list, same material branch, same grid cells); with longer jobs the lanes of a warp drift
// apart over the frame.  Measured (profiles/logs/r02w_*, r02x_*): config 4 1625 -> 1542 ms, config 5 1323 -> 1257 ms, config 2
// 39.8 -> 36.6 ms against 8-30 samples per job -- and the drain at the end of the launch is one sample long.
struct JobPlan {
    unsigned long long pix_local;      // pixels rendered by this launch
    unsigned long long total_jobs;
    unsigned long long jobs_a;         // jobs of the full bands; the jobs of the last (partial) band follow
    unsigned long long band_jobs;      // band_pix * chunks
    unsigned long long rp_b;           // first reversed pixel of the last band = (bands - 1) * band_pix
    unsigned long long magic_band_jobs, magic_width;      // floor(2^64 / d) + 1 (0 encodes d == 1)
    unsigned long long magic_pix[2];                      // ... for band_pix / last_pix
    unsigned int band_pix, last_pix;   // pixels per full band / in the last band
    unsigned int tiles_per_row;        // > 0: pixels are numbered in 8 x 4 tiles (width % 8 == 0, rows % 4 == 0), else row-major
    unsigned long long magic_tpr;
    int chunks;                        // sample ranges per pixel = ceil(s_count / spj)
    int spj;                           // samples per job
    int s_begin, s_count;              // this launch traces samples [s_begin, s_begin + s_count) of every pixel
};

template <typename T> struct TraceArgs {
    DevCamera<T> cam;
    SceneBlob scene;
    PhiloxKeys keys;                   // Philox round keys of the seed
    int max_depth;
    int width, height;
    int tile_rows, rank, world;        // local row -> global row (world == 1: identity)
    JobPlan plan;
    long long *acc;                    // [pix_local][3] fixed-point radiance sums (rt_device.cuh: accumulate)
    unsigned long long *queue;         // [0] job cursor, [1] segments, [2] paths, [3] BVH nodes / grid cells, [4] exact sphere tests,
                                       // [5] binned camera segments, [6] filter tests
    BvhView bvh;                       // RT_ACCEL_LBVH only
    int bvh_steps;                     // at most this many node visits per loop turn ...
    int bvh_min_active;                // ... and the round ends once fewer lanes than this are still traversing
    const unsigned int *bins;          // trace_kernel_pb only: per-tile candidate lists of the camera rays (rt_primary_bins.cuh)
    int tiles_x, tiles;
    int pb_rounds, pb_min;             // camera-ray rounds per loop turn; rounds after the first need this many fresh lanes
    int pb_cohort;                     // lanes out of work claim new jobs only when this many can claim together (1: at once)
};

// ------------------------------------------------------------------------------------------
template <typename T> struct PathState {
    Vec3<T> o, d, att;
    T puy;                // y of the normalised PRIMARY direction (sky term, GF camera.h:121)
};

// get_ray (GF camera.h:145-155) with Philox dimension 0 of (pixel, sample): block 0 = (x jitter, y jitter, first lens
// candidate), every later block two lens candidates (float; double: one uniform per two words, so block 0 = the jitter and
// every later block one candidate).
// `ph` is opened on (pixel, sample, 0).  HAVE0 (what every kernel uses): it already holds block 0, computed by the caller
// with all lanes converged; HAVE0 = false computes it at the one Philox site of the loop below -- smaller code, but measured
// 5 % SLOWER in all three trace kernels (profiles/logs/r02i_philox_site_ab.log).
template <typename T, bool HAVE0 = true>
__device__ __forceinline__ void camera_ray(const TraceArgs<T> &A, int i, int j, Philox &ph, PathState<T> &ps) {
    using N = Num<T>;
    const bool lens = !(A.cam.defocus_angle <= T(0));
    Vec3<T> target;
    target.x = target.y = target.z = T(0);
    T q0 = T(0), q1 = T(0);
    // defocus_disk_sample / random_in_unit_disk (GF camera.h:73-76, vec3.h:109-115): first candidate inside the unit disk
    for (uint32_t b = 0;; ++b) {
        if (!HAVE0 || b > 0) ph.block(b);
        bool done = false;
        if (b == 0) {
            T ux, uy;
            if (N::words_per_uniform == 1) { ux = N::uniform(ph.w[0], 0); uy = N::uniform(ph.w[1], 0); }
            else { ux = N::uniform(ph.w[0], ph.w[1]); uy = N::uniform(ph.w[2], ph.w[3]); }
            const T px = N::add(static_cast<T>(i), N::sub(ux, T(0.5)));
            const T py = N::add(static_cast<T>(j), N::sub(uy, T(0.5)));
            target.x = N::fma(py, A.cam.dv.x, N::fma(px, A.cam.du.x, A.cam.pixel00.x));
            target.y = N::fma(py, A.cam.dv.y, N::fma(px, A.cam.du.y, A.cam.pixel00.y));
            target.z = N::fma(py, A.cam.dv.z, N::fma(px, A.cam.du.z, A.cam.pixel00.z));
            if (!lens) break;
        }
        if (N::words_per_uniform == 1) {
            if (b > 0) {                                     // words 0, 1 of blocks 1, 2, ...
                q0 = N::fma(N::uniform(ph.w[0], 0), T(2), T(-1));
                q1 = N::fma(N::uniform(ph.w[1], 0), T(2), T(-1));
                done = N::fma(q1, q1, N::mul(q0, q0)) < T(1);
            }
            if (!done) {                                     // words 2, 3 of every block
                q0 = N::fma(N::uniform(ph.w[2], 0), T(2), T(-1));
                q1 = N::fma(N::uniform(ph.w[3], 0), T(2), T(-1));
                done = N::fma(q1, q1, N::mul(q0, q0)) < T(1);
            }
        } else if (b > 0) {
            q0 = N::fma(N::uniform(ph.w[0], ph.w[1]), T(2), T(-1));
            q1 = N::fma(N::uniform(ph.w[2], ph.w[3]), T(2), T(-1));
            done = N::fma(q1, q1, N::mul(q0, q0)) < T(1);
        }
        if (done) break;
    }
    Vec3<T> o = A.cam.center;
    if (lens) {
        o.x = N::fma(q1, A.cam.disk_v.x, N::fma(q0, A.cam.disk_u.x, A.cam.center.x));
        o.y = N::fma(q1, A.cam.disk_v.y, N::fma(q0, A.cam.disk_u.y, A.cam.center.y));
        o.z = N::fma(q1, A.cam.disk_v.z, N::fma(q0, A.cam.disk_u.z, A.cam.center.z));
    }
    ps.o = o;
    ps.d.x = N::sub(target.x, o.x); ps.d.y = N::sub(target.y, o.y); ps.d.z = N::sub(target.z, o.z);
    ps.att.x = ps.att.y = ps.att.z = T(1);
    const T len = N::sqrt(dot3(ps.d, ps.d));
    ps.puy = N::mul(N::rcp(len), ps.d.y);
}

// background of the primary ray (GF camera.h:120-124): a = 0.5*(uy + 1.0) in double
template <typename T> __device__ __forceinline__ void sky(T puy, T &r, T &g, T &b);
template <> __device__ __forceinline__ void sky<float>(float puy, float &r, float &g, float &b) {
    const double a = __dmul_rn(0.5, __dadd_rn((double)puy, 1.0));
    const float t1 = __double2float_rn(__dsub_rn(1.0, a));
    const float af = __double2float_rn(a);
    r = __fmaf_rn(af, 0.5f, t1);
    g = __fmaf_rn(af, 0.7f, t1);
    b = __fadd_rn(af, t1);
}
template <> __device__ __forceinline__ void sky<double>(double puy, double &r, double &g, double &b) {
    const double a = __dmul_rn(0.5, __dadd_rn(puy, 1.0));
    const double t1 = __dsub_rn(1.0, a);
    r = __fma_rn(a, 0.5, t1);
    g = __fma_rn(a, 0.7, t1);
    b = __dadd_rn(a, t1);
}

// The point random_unit_vector (GF vec3.h:117-127) tries with one Philox block, and whether it accepts it
// (unit_min < |c|^2 <= 1); float only -- a double candidate takes two blocks (scatter's own loop).
__device__ __forceinline__ bool ball_candidate(const Philox &ph, Vec3<float> &c) {
    using N = Num<float>;
    c.x = N::fma(N::uniform(ph.w[0], 0), 2.0f, -1.0f);
    c.y = N::fma(N::uniform(ph.w[1], 0), 2.0f, -1.0f);
    c.z = N::fma(N::uniform(ph.w[2], 0), 2.0f, -1.0f);
    const float l2 = dot3(c, c);
    return N::unit_min() < l2 && l2 <= 1.0f;
}

// Warp-cooperative tail of random_unit_vector's rejection loop (float kernels; call with all 32 lanes).  A lane whose first
// candidate (block 0 of its own (pixel, sample, dimension) counter) was rejected enters with need = true; the accepted
// candidate is the one of the SMALLEST block index that passes, exactly as the sequential loop finds it.  Left to itself the
// loop runs at 6-8 active threads (half of the searching lanes drop out per trip, the warp waits for the unluckiest: 11 % of the
// grid kernel's instructions); here every lane of the warp -- searching, shading a dielectric, or idle -- computes one block per
// trip for the first searching lane at or above it: lane f is served by the lanes (previous searching lane, f], helper L tries
// block next_f + (f - L), and f takes the passing helper closest to itself.  Two trips on average instead of four to five.
__device__ __forceinline__ void coop_unit_vector(const PhiloxKeys &keys, bool need, uint32_t c0, uint32_t c1, uint32_t c2, Vec3<float> &c) {
    const int lane = threadIdx.x & 31;
    uint32_t next = 1u;
    unsigned searching = __ballot_sync(FULL, need);
    while (searching) {
        const unsigned up = searching >> lane;                    // bit 0 = this lane
        const int dist = up ? __ffs(up) - 1 : 0;
        const int f = lane + dist;
        const uint32_t h0 = __shfl_sync(FULL, c0, f), h1 = __shfl_sync(FULL, c1, f), h2 = __shfl_sync(FULL, c2, f);
        const uint32_t hk = __shfl_sync(FULL, next, f) + (uint32_t)dist;
        Vec3<float> hc;
        hc.x = hc.y = hc.z = 0.0f;
        bool ok = false;
        if (up) {
            Philox ph;
            ph.open(keys, h0, h1, h2);
            ph.block(hk);
            ok = ball_candidate(ph, hc);
        }
        const unsigned passed = __ballot_sync(FULL, ok);
        int src = lane;
        bool found = false;
        if (need) {
            const unsigned below = searching & ((1u << lane) - 1u);
            const int prev = below ? 31 - __clz(below) : -1;      // the previous searching lane
            const unsigned upto_me = 0xffffffffu >> (31 - lane), upto_prev = prev >= 0 ? 0xffffffffu >> (31 - prev) : 0u;
            const unsigned mine = passed & upto_me & ~upto_prev;
            if (mine) { src = 31 - __clz(mine); found = true; need = false; }
            else next += (uint32_t)(lane - prev);
        }
        const float sx = __shfl_sync(FULL, hc.x, src), sy = __shfl_sync(FULL, hc.y, src), sz = __shfl_sync(FULL, hc.z, src);
        if (found) { c.x = sx; c.y = sy; c.z = sz; }
        searching = __ballot_sync(FULL, need);
    }
}

// One bounce: hit record (GF hittable.h:58-63) + the material switch of GF camera.h:92-108, given the random numbers of the
// bounce: `schlick_u` for a dielectric, the accepted candidate `cand` of random_unit_vector (not yet normalised) otherwise.
// Returns false when the path is absorbed (metal scattered below the surface, GF material.h:58).
template <typename T>
__device__ __forceinline__ bool scatter_with(const SceneView<T> &sc, const Hit<T> &hit, int type, T schlick_u, const Vec3<T> &cand,
                                             PathState<T> &ps) {
    using N = Num<T>;
    RT_CHECK(hit.id >= 0 && hit.id < sc.n, 302);
    const typename N::vec4 s = sc.geom[hit.id];
    const typename N::vec4 m = sc.matl[hit.id];
    const Vec3<T> o = ps.o, d = ps.d;
    Vec3<T> p;
    p.x = N::fma(hit.t, d.x, o.x); p.y = N::fma(hit.t, d.y, o.y); p.z = N::fma(hit.t, d.z, o.z);
    const T inv_r = sc.rinv[hit.id];                            // 1/radius, rounded once on the host (== rcp.rn)
    Vec3<T> n;
    n.x = N::mul(N::sub(p.x, s.x), inv_r); n.y = N::mul(N::sub(p.y, s.y), inv_r); n.z = N::mul(N::sub(p.z, s.z), inv_r);
    const bool front = dot3(d, n) < T(0);
    if (!front) { n.x = -n.x; n.y = -n.y; n.z = -n.z; }

    Vec3<T> nd;
    if (type == RT_DIELECTRIC) {
        // dieletric_scatter (GF material.h:68-89), reflect/refract (GF vec3.h:129-138)
        const T ri = front ? N::rcp(m.w) : m.w;
        const T inv = N::rcp(N::sqrt(dot3(d, d)));
        Vec3<T> ud;
        ud.x = N::mul(inv, d.x); ud.y = N::mul(inv, d.y); ud.z = N::mul(inv, d.z);
        Vec3<T> nud;
        nud.x = -ud.x; nud.y = -ud.y; nud.z = -ud.z;
        const T cos_t = N::min(dot3(nud, n), T(1));
        const T sin_t = N::sqrt(N::fma(-cos_t, cos_t, T(1)));
        bool reflect = N::mul(ri, sin_t) > T(1);
        if (!reflect) {
            // Schlick (GF material.h:62-66); (1-cos)^5 by repeated multiplication
            T r0 = N::div(N::sub(T(1), ri), N::add(T(1), ri));
            r0 = N::mul(r0, r0);
            const T x1 = N::sub(T(1), cos_t), x2 = N::mul(x1, x1), x4 = N::mul(x2, x2), x5 = N::mul(x4, x1);
            const T refl = N::fma(N::sub(T(1), r0), x5, r0);
            reflect = refl > schlick_u;
        }
        if (reflect) {
            const T k = N::mul(T(2), dot3(ud, n));
            nd.x = N::fma(-k, n.x, ud.x); nd.y = N::fma(-k, n.y, ud.y); nd.z = N::fma(-k, n.z, ud.z);
        } else {
            Vec3<T> perp;
            perp.x = N::mul(ri, N::fma(cos_t, n.x, ud.x));
            perp.y = N::mul(ri, N::fma(cos_t, n.y, ud.y));
            perp.z = N::mul(ri, N::fma(cos_t, n.z, ud.z));
            const T k = -N::sqrt(N::abs(N::sub(T(1), dot3(perp, perp))));
            nd.x = N::fma(k, n.x, perp.x); nd.y = N::fma(k, n.y, perp.y); nd.z = N::fma(k, n.z, perp.z);
        }
    } else {
        Vec3<T> uv;
        {
            const T inv = N::rcp(N::sqrt(dot3(cand, cand)));      // unit_vector of the accepted candidate (GF vec3.h:126)
            uv.x = N::mul(inv, cand.x); uv.y = N::mul(inv, cand.y); uv.z = N::mul(inv, cand.z);
        }
        if (type == RT_LAMBERTIAN) {
            // lambertian_scatter (GF material.h:38-49)
            nd.x = N::add(n.x, uv.x); nd.y = N::add(n.y, uv.y); nd.z = N::add(n.z, uv.z);
            if (N::abs(nd.x) < N::near_zero() && N::abs(nd.y) < N::near_zero() && N::abs(nd.z) < N::near_zero()) nd = n;
        } else {
            // metal_scatter (GF material.h:51-59)
            const T k = N::mul(T(2), dot3(d, n));
            Vec3<T> rf;
            rf.x = N::fma(-k, n.x, d.x); rf.y = N::fma(-k, n.y, d.y); rf.z = N::fma(-k, n.z, d.z);
            const T inv = N::rcp(N::sqrt(dot3(rf, rf)));
            nd.x = N::fma(m.w, uv.x, N::mul(inv, rf.x));
            nd.y = N::fma(m.w, uv.y, N::mul(inv, rf.y));
            nd.z = N::fma(m.w, uv.z, N::mul(inv, rf.z));
            if (!(dot3(nd, n) > T(0))) return false;
        }
        ps.att.x = N::mul(ps.att.x, m.x); ps.att.y = N::mul(ps.att.y, m.y); ps.att.z = N::mul(ps.att.z, m.z);
    }
    ps.o = p;
    ps.d = nd;
    return true;
}

// scatter with the sequential draw.  `ph` is opened on (pixel, sample, depth+1): block 0 feeds the Schlick uniform of a
// dielectric (word 0; double: words 0, 1) or the first random_unit_vector candidate (one candidate per block -- float -- or
// per two blocks -- double), later blocks the later candidates.  HAVE0: `ph` already holds block 0; otherwise it is computed
// at the one Philox site of the candidate loop (smaller, but measured 5 % slower).
template <typename T, bool HAVE0 = true>
__device__ __forceinline__ bool scatter(const SceneView<T> &sc, const Hit<T> &hit, Philox &ph, PathState<T> &ps) {
    using N = Num<T>;
    RT_CHECK(hit.id >= 0 && hit.id < sc.n, 302);
    const int type = sc.type[hit.id];
    Vec3<T> cand;
    cand.x = cand.y = cand.z = T(1);
    T schlick_u = T(0);
    for (uint32_t k = 0;; ++k) {
        T ux, uy, uz;
        if (N::words_per_uniform == 1) {
            if (!HAVE0 || k > 0) ph.block(k);
            if (type == RT_DIELECTRIC) { schlick_u = N::uniform(ph.w[0], 0); break; }
            ux = N::uniform(ph.w[0], 0); uy = N::uniform(ph.w[1], 0); uz = N::uniform(ph.w[2], 0);
        } else {
            if (!HAVE0 || k > 0) ph.block(2u * k);
            if (type == RT_DIELECTRIC) { schlick_u = N::uniform(ph.w[0], ph.w[1]); break; }
            ux = N::uniform(ph.w[0], ph.w[1]); uy = N::uniform(ph.w[2], ph.w[3]);
            ph.block(2u * k + 1u);
            uz = N::uniform(ph.w[0], ph.w[1]);
        }
        cand.x = N::fma(ux, T(2), T(-1)); cand.y = N::fma(uy, T(2), T(-1)); cand.z = N::fma(uz, T(2), T(-1));
        const T l2 = dot3(cand, cand);
        if (N::unit_min() < l2 && l2 <= T(1)) break;
    }
    return scatter_with<T>(sc, hit, type, schlick_u, cand, ps);
}

template <typename T>
__device__ __forceinline__ int global_row(const TraceArgs<T> &A, int local_row) {
    if (A.world == 1) return local_row;
    const int tile = local_row / A.tile_rows, within = local_row - tile * A.tile_rows;
    return (tile * A.world + A.rank) * A.tile_rows + within;
}

// decode_job turns a queue index into the pixel and the sample range (JobPlan above).  Exact divisions by multiply-high:
// the host checks that every dividend * divisor stays below 2^63.
struct JobInfo {
    int pi, pj, sample, sample_end;
    uint32_t pixel;                    // global pixel index (Philox counter word 0)
    uint32_t local;                    // index of the pixel in this launch's accumulators
};
__device__ __forceinline__ unsigned long long div_magic(unsigned long long x, unsigned long long magic) {
    return magic ? __umul64hi(x, magic) : x;
}
template <typename T>
__device__ __forceinline__ JobInfo decode_job(const TraceArgs<T> &A, unsigned long long job) {
    const JobPlan &P = A.plan;
    JobInfo J;
    const int reg = job >= P.jobs_a;                                   // the last band (the top rows; partial)
    unsigned long long r, rp0;
    if (reg) { r = job - P.jobs_a; rp0 = P.rp_b; }
    else {
        const unsigned long long band = div_magic(job, P.magic_band_jobs);
        r = job - band * P.band_jobs;
        rp0 = band * P.band_pix;
    }
    const unsigned long long bp = reg ? P.last_pix : P.band_pix;
    const unsigned long long cl = div_magic(r, P.magic_pix[reg]);
    const unsigned long long lp = P.pix_local - 1ull - (rp0 + (r - cl * bp));   // bands are counted from the bottom of the frame
    int lr;
    if (P.tiles_per_row) {
        // 8 x 4 pixel tiles, row-major inside a tile and across tiles: the lanes that claim together cover a compact block
        // (one tile list, neighbouring rays) instead of a strip of a row
        const unsigned long long tile = lp >> 5;
        const unsigned int w = (unsigned int)lp & 31u;
        const unsigned long long ty = div_magic(tile, P.magic_tpr);
        J.pi = (int)((tile - ty * P.tiles_per_row) * 8ull) + (int)(w & 7u);
        lr = (int)(ty * 4ull) + (int)(w >> 3);
    } else {
        lr = (int)div_magic(lp, P.magic_width);
        J.pi = (int)(lp - (unsigned long long)lr * A.width);
    }
    J.pj = global_row(A, lr);
    J.pixel = (uint32_t)J.pj * (uint32_t)A.width + (uint32_t)J.pi;
    J.local = (uint32_t)lr * (uint32_t)A.width + (uint32_t)J.pi;
    J.sample = P.s_begin + (int)cl * P.spj;
    J.sample_end = min(J.sample + P.spj, P.s_begin + P.s_count);
    RT_CHECK(job < P.total_jobs && lp < P.pix_local && J.local < P.pix_local && cl < (unsigned long long)P.chunks && r - cl * bp < bp, 201);
    RT_CHECK(J.pi >= 0 && J.pi < A.width && J.pj >= 0 && J.pj < A.height, 202);
    RT_CHECK(J.sample >= P.s_begin && J.sample < J.sample_end && J.sample_end <= P.s_begin + P.s_count, 203);
    return J;
}

// One atomic per warp for all lanes that ran out of work (`want` = ballot of those lanes): this lane's queue index.
template <typename T>
__device__ __forceinline__ unsigned long long claim_job(const TraceArgs<T> &A, int lane, unsigned want) {
    const int leader = __ffs(want) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(A.queue, (unsigned long long)__popc(want));
    base = __shfl_sync(FULL, base, leader);
    return base + (unsigned long long)__popc(want & ((1u << lane) - 1u));
}

// per-warp reduction of the lanes' work counters into the queue words 1..6: two REDUX per counter (low and high halves, so
// the 32-lane sum cannot overflow) instead of five 64-bit shuffle rounds -- the kernel's code footprint matters (I-cache)
__device__ __forceinline__ void flush_counters(unsigned long long *queue, int lane, unsigned n_seg, unsigned n_path, unsigned n_nodes,
                                               unsigned n_exact, unsigned n_binned, unsigned n_filt) {
    const unsigned v[6] = {n_seg, n_path, n_nodes, n_exact, n_binned, n_filt};
#pragma unroll 1
    for (int q = 0; q < 6; ++q) {
        const unsigned lo = __reduce_add_sync(FULL, v[q] & 0xffffu), hi = __reduce_add_sync(FULL, v[q] >> 16);
        const unsigned long long sum = (unsigned long long)lo + ((unsigned long long)hi << 16);
        if (lane == 0 && sum) atomicAdd(queue + 1 + q, sum);
    }
}

template <typename T, int ACCEL>
__global__ void __launch_bounds__(TRACE_BLOCK, sizeof(T) == 4 ? RT_TRACE_MIN_BLOCKS : 2) trace_kernel(const __grid_constant__ TraceArgs<T> A) {
    using N = Num<T>;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    SceneView<T> sc;
    unsigned short *cand = nullptr;
    ScanGeom geo = scan_geom(0u, 0);
    if (ACCEL == RT_ACCEL_LINEAR) {
        // whole scene -> shared memory, one TMA bulk copy per CTA
        stage_scene(smem, A.scene.base, A.scene.bytes, &bar);
        sc = view_of<T>(smem, A.scene);
        cand = reinterpret_cast<unsigned short *>(smem + A.scene.bytes) + threadIdx.x;
        geo = scan_geom(smem_u32(smem), A.scene);
    } else {
        // large scenes: geometry through the LBVH in global memory (L1/L2), materials by slot
        sc = view_of<T>(A.scene.base, A.scene);
    }
    unsigned int n_nodes = 0, n_tests = 0;
    constexpr bool LB = ((ACCEL == RT_ACCEL_LBVH || ACCEL == ACCEL_LBVH_COMPACT) && sizeof(T) == 4);
    constexpr bool RAYD = (ACCEL == ACCEL_LBVH_COMPACT);
    BvhTrav tv;                                   // LBVH only: resumable traversal (stack in local memory)
    BvhStack bvh_stack;
    tv.node = -1;

    const int lane = threadIdx.x & 31;
    enum { NEED_JOB = 0, ACTIVE = 1, DEAD = 2 };
    int state = NEED_JOB;
    bool fresh = false;                 // this lane starts a new sample at the top of the next turn
    PathState<T> ps;
    ps.o = A.cam.center;
    ps.d.x = T(0); ps.d.y = T(1); ps.d.z = T(0);
    ps.att.x = ps.att.y = ps.att.z = T(1);
    ps.puy = T(0);
    Hit<T> hit;                         // pending hit of the previous turn's scan
    hit.t = N::inf();
    hit.id = -1;
    int pi = 0, pj = 0, sample = 0, sample_end = 0, depth = 0;
    uint32_t pixel = 0, local = 0;
    unsigned int n_seg = 0, n_path = 0;
    ScanCount cnt{0u, 0u};

    // a path ended with radiance (cr,cg,cb): add it to the pixel's accumulators (GF camera.h:160), move to the next
    // sample of the job or ask for the next job
    auto end_path = [&](T cr, T cg, T cb) {
        RT_CHECK(local < A.plan.pix_local && state == ACTIVE, 301);
        accumulate<T>(A.acc, local, cr, cg, cb);
        ++n_path;
        if (++sample == sample_end) state = NEED_JOB;
        else fresh = true;
    };
    auto end_black = [&]() {
        ++n_path;
        if (++sample == sample_end) state = NEED_JOB;
        else fresh = true;
    };

    for (;;) {
        // ---- job fetch: one atomic per warp for all lanes that ran out of work ----
        const unsigned want = __ballot_sync(FULL, state == NEED_JOB);
        if (want) {
            const unsigned long long claimed = claim_job(A, lane, want);
            if (state == NEED_JOB) {
                if (claimed < A.plan.total_jobs) {
                    const JobInfo J = decode_job(A, claimed);
                    pi = J.pi; pj = J.pj; pixel = J.pixel; local = J.local;
                    sample = J.sample; sample_end = J.sample_end;
                    state = ACTIVE;
                    fresh = true;
                } else {
                    state = DEAD;
                }
            }
        }
        if (__all_sync(FULL, state == DEAD)) break;

        // LBVH: a lane whose traversal is still in flight skips shading/regeneration this turn
        const bool idle = (tv.node < 0);                                  // always true for the linear scan
        auto ready = [&]() { return LB ? (idle && state == ACTIVE) : (state == ACTIVE); };

        // ---- one Philox block per lane and turn, shared by the two consumers: dimension 0 feeds
        //      the camera ray of a fresh sample, dimension depth+1 the scatter of the pending hit ----
        Philox ph;
        ph.open(A.keys, pixel, (uint32_t)sample, fresh ? 0u : (uint32_t)(depth + 1));
        ph.block(0);
        if (ready() && !fresh) {
            // shade the hit found by the previous scan
            const bool alive = scatter(sc, hit, ph, ps);
            if (!alive || ++depth >= A.max_depth) {                       // GF camera.h:117 / :84,127 -> black
                end_black();
                if (state == ACTIVE) {                                    // rare: regenerate right away
                    ph.open(A.keys, pixel, (uint32_t)sample, 0u);
                    ph.block(0);
                }
            }
        }
        // ---- path regeneration: a finished lane starts its next sample in place ----
        if (ready() && fresh) {
            camera_ray(A, pi, pj, ph, ps);
            depth = 0;
            fresh = false;
        }

        bool landed;                                  // this lane's closest hit became available this turn
        if constexpr (LB) {
            // ---- LBVH: start the traversal of the new ray, then a bounded number of node visits for
            //      every lane with a traversal in flight ----
            const bool launched = ready();
            if (launched) bvh_start<RAYD>(A.bvh, ps.o, ps.d, tv, n_tests);
            const bool flying = launched || (state == ACTIVE && !idle);
#pragma unroll 1
            for (int step = 0; step < A.bvh_steps; ++step) {
                // stop early when enough lanes have finished: they are shaded and regenerated together,
                // and the node loop never runs with a nearly empty warp
                const int flying_lanes = __popc(__ballot_sync(FULL, tv.node >= 0));
                if (flying_lanes == 0 || (flying_lanes < A.bvh_min_active && step > 0)) break;
                if (tv.node >= 0) bvh_step<RAYD>(A.bvh, ps.o, ps.d, tv, bvh_stack, n_nodes, n_tests);
            }
            landed = flying && tv.node < 0;
            if (landed) hit = tv.hit;
        } else {
            // ---- closest hit over all slots (all 32 lanes, uniform trip count) ----
            hit = closest_hit<T>(geo, A.scene.n, ps.o, ps.d, cand, TRACE_BLOCK, cnt);
            landed = (state == ACTIVE);
        }

        // ---- misses end the path here (no random numbers needed); hits are shaded next turn ----
        if (landed) {
            ++n_seg;
            if (hit.id < 0) {
                T sr, sg, sb;
                sky<T>(ps.puy, sr, sg, sb);
                end_path(N::mul(ps.att.x, sr), N::mul(ps.att.y, sg), N::mul(ps.att.z, sb));
            }
        }
    }

    // ---- work counters (reference-equivalent work: segments x slots; executed work: filter + exact tests, node visits) ----
    flush_counters(A.queue, lane, n_seg, n_path, n_nodes, n_tests + cnt.exact, 0u, cnt.filt);
}

}  // namespace rt

#include "rt_wavefront.cuh"
#include "rt_primary_bins.cuh"

namespace rt {

// ------------------------------------------------------------------------------------------
// out[p] = gamma(scale * float(sum_g acc_g[p] * 2^-40)); 4 pixels per thread, 16-byte loads and stores.
// `src` lists the accumulation buffers to add: one for a whole-frame or rows render, one per GPU for the spp split --
// there the pointers are the PEERS' buffers (NVLink P2P loads, rt_enable_peer_access), so the cross-GPU reduction, the
// scale, the gamma and the store are one kernel; integer sums make the result independent of the order.
struct RowPlacement { int width, tile_rows, rank, world; };    // world == 1: rows stay where they are
constexpr int ACC_SOURCES_MAX = 16;
struct AccSources { const long long *p[ACC_SOURCES_MAX]; int n; };

__device__ __forceinline__ unsigned long long placed_pixel(const RowPlacement &rp, unsigned long long p) {
    if (rp.world == 1) return p;
    const int lr = (int)(p / (unsigned long long)rp.width), col = (int)(p - (unsigned long long)lr * rp.width);
    const int tile = lr / rp.tile_rows, within = lr - tile * rp.tile_rows;
    return (unsigned long long)((tile * rp.world + rp.rank) * rp.tile_rows + within) * rp.width + col;
}

template <typename T> __device__ __forceinline__ T unfix(long long v);
template <> __device__ __forceinline__ float unfix<float>(long long v) {
    return __double2float_rn(__dmul_rn(__ll2double_rn(v), 9.094947017729282e-13));       // 2^-40
}
template <> __device__ __forceinline__ double unfix<double>(long long v) { return __dmul_rn(__ll2double_rn(v), 9.094947017729282e-13); }

template <typename T>
__global__ void __launch_bounds__(256) finalize_kernel(const __grid_constant__ AccSources src, unsigned long long pix, T scale,
                                                       T *__restrict__ out, const RowPlacement rp) {
    using N = Num<T>;
    const unsigned long long quad = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long p0 = quad * 4ull;
    if (p0 >= pix) return;
    T v[12];
    const int cnt = (pix - p0) < 4ull ? (int)(pix - p0) : 4;
    long long sum[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) sum[k] = 0;
    for (int g = 0; g < src.n; ++g) {
        const long long *a = src.p[g] + 3ull * p0;                      // 96-byte groups: 16-byte aligned
        if (cnt == 4) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const longlong2 q = *reinterpret_cast<const longlong2 *>(a + 2 * k);
                sum[2 * k] += q.x; sum[2 * k + 1] += q.y;
            }
        } else {
            for (int k = 0; k < 3 * cnt; ++k) sum[k] += a[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        T c = N::mul(unfix<T>(sum[k]), scale);                          // GF camera.h:167
        v[k] = c > T(0) ? N::sqrt(c) : T(0);                            // GF color.h:10-13
    }
    // row placement (multi-GPU direct stores into the full frame): a group of 4 pixels stays inside one
    // row when the width is a multiple of 4, otherwise fall back to per-pixel stores
    if (rp.world > 1 && (rp.width & 3)) {
        for (int k = 0; k < cnt; ++k) {
            T *d1 = out + placed_pixel(rp, p0 + k) * 3ull;
            d1[0] = v[3 * k]; d1[1] = v[3 * k + 1]; d1[2] = v[3 * k + 2];
        }
        return;
    }
    T *dst = out + placed_pixel(rp, p0) * 3ull;
    if (cnt == 4 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        constexpr int per = 16 / sizeof(T);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        const uint4 *s4 = reinterpret_cast<const uint4 *>(v);
#pragma unroll
        for (int k = 0; k < 12 / per; ++k) d4[k] = s4[k];
    } else {
        for (int k = 0; k < 3 * cnt; ++k) dst[k] = v[k];
    }
}

// The same conversion when the rows stay where they are (one GPU, the spp split, rt_finalize[_sum]): value c of the frame
// depends on accumulator c alone, so the frame is converted as a flat array -- lane i of a warp reads the i-th 16-byte pair of
// sums and writes the i-th pair of values.  finalize_kernel's 96-byte groups per thread put the 16-byte loads of a warp 96 bytes
// apart (every load instruction touches 32 sectors for 512 useful bytes): 0.108 ms = 2.8 TB/s for a 4K frame; here every load
// and store instruction covers whole lines.  HBM-bound: 24 B read per pixel and source + 12 B written.
constexpr int FINALIZE_FLAT_UNROLL = 4;
template <typename T> struct Pair2;
template <> struct Pair2<float> { using type = float2; };
template <> struct Pair2<double> { using type = double2; };
template <typename T>
__device__ __forceinline__ T encode_value(long long sum, T scale) {
    using N = Num<T>;
    const T c = N::mul(unfix<T>(sum), scale);                              // GF camera.h:167
    return c > T(0) ? N::sqrt(c) : T(0);                                    // GF color.h:10-13
}
template <typename T>
__global__ void __launch_bounds__(256) finalize_flat_kernel(const __grid_constant__ AccSources src, unsigned long long vals, T scale,
                                                            T *__restrict__ out) {
    constexpr int U = FINALIZE_FLAT_UNROLL;
    const unsigned long long pairs = vals >> 1;
    const unsigned long long base = (unsigned long long)blockIdx.x * (U * 256ull) + threadIdx.x;
    long long sx[U], sy[U];
#pragma unroll
    for (int u = 0; u < U; ++u) sx[u] = sy[u] = 0;
    for (int g = 0; g < src.n; ++g) {
        const int4 *a = reinterpret_cast<const int4 *>(src.p[g]);              // one 16-byte load per pair of sums
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned long long j = base + (unsigned long long)u * 256ull;
            if (j < pairs) {
                const int4 q = __ldg(a + j);
                sx[u] += (long long)(((unsigned long long)(unsigned int)q.y << 32) | (unsigned int)q.x);
                sy[u] += (long long)(((unsigned long long)(unsigned int)q.w << 32) | (unsigned int)q.z);
            }
        }
    }
    typename Pair2<T>::type *o2 = reinterpret_cast<typename Pair2<T>::type *>(out);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const unsigned long long j = base + (unsigned long long)u * 256ull;
        if (j < pairs) {
            typename Pair2<T>::type v;
            v.x = encode_value<T>(sx[u], scale);
            v.y = encode_value<T>(sy[u], scale);
            o2[j] = v;
        }
    }
    if ((vals & 1ull) && blockIdx.x == 0 && threadIdx.x == 0) {                 // an odd number of pixels: the last value
        long long t = 0;
        for (int g = 0; g < src.n; ++g) t += src.p[g][vals - 1];
        out[vals - 1] = encode_value<T>(t, scale);
    }
}

// ------------------------------------------------------------------------------------------
template <typename T, int ACCEL>
__global__ void __launch_bounds__(TRACE_BLOCK) primary_kernel(const __grid_constant__ DevCamera<T> cam,
                                                              const __grid_constant__ SceneBlob scene,
                                                              const __grid_constant__ BvhView bvh, int width,
                                                              int height, int *__restrict__ ids, T *__restrict__ ts) {
    using N = Num<T>;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    unsigned short *cand = nullptr;
    ScanGeom geo = scan_geom(0u, 0);
    if (ACCEL == RT_ACCEL_LINEAR) {
        stage_scene(smem, scene.base, scene.bytes, &bar);
        cand = reinterpret_cast<unsigned short *>(smem + scene.bytes) + threadIdx.x;
        geo = scan_geom(smem_u32(smem), scene);
    }
    unsigned int n_nodes = 0, n_tests = 0;
    const long long npix = (long long)width * height;
    // the closest-hit scan is warp-cooperative: every lane of a warp takes every trip, lanes past the end
    // trace pixel 0 and store nothing
    for (long long k0 = (long long)blockIdx.x * blockDim.x; k0 < npix; k0 += (long long)gridDim.x * blockDim.x) {
        const bool valid = k0 + threadIdx.x < npix;
        const long long k = valid ? k0 + threadIdx.x : 0;
        const int j = (int)(k / width), i = (int)(k - (long long)j * width);
        const T fi = static_cast<T>(i), fj = static_cast<T>(j);
        Vec3<T> d;
        d.x = N::sub(N::fma(fj, cam.dv.x, N::fma(fi, cam.du.x, cam.pixel00.x)), cam.center.x);
        d.y = N::sub(N::fma(fj, cam.dv.y, N::fma(fi, cam.du.y, cam.pixel00.y)), cam.center.y);
        d.z = N::sub(N::fma(fj, cam.dv.z, N::fma(fi, cam.du.z, cam.pixel00.z)), cam.center.z);
        Hit<T> hit;
        if constexpr (ACCEL == RT_ACCEL_LBVH && sizeof(T) == 4) hit = bvh_closest_hit(bvh, cam.center, d, n_nodes, n_tests);
        else if constexpr (ACCEL == RT_ACCEL_GRID)
            hit = grid_closest_hit<T>(g_grid, static_cast<const typename Num<T>::vec4 *>(scene.base), cam.center, d, n_nodes, n_tests);
        else hit = closest_hit<T>(geo, scene.n, cam.center, d, cand, TRACE_BLOCK);
        if (valid) {
            ids[k] = hit.id;
            ts[k] = hit.t;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Audit of the paired scan's conservative filter (rt_filter_audit): every generated ray is tested
// against every filtered slot with BOTH the reference's exact discriminant (disc_of) and the filter
// (filter_ray / filter_value, the code the scan runs).  out[0] pairs, out[1] exact passes (disc >= 0),
// out[2] filter passes, out[3] MISSES (exact pass the filter rejected: must be 0), out[4] rays skipped
// as degenerate.  Ray kinds, by index: 0 jittered camera rays; 1 bounce-like rays (origin on a random
// sphere, random direction); 2 / 3 rays aimed at the silhouette of a random sphere from the camera /
// from a point on another sphere, nudged by a few ulp -- the discriminant's sign is decided by rounding.
__global__ void __launch_bounds__(256) filter_audit_kernel(const __grid_constant__ SceneBlob scene,
                                                           const __grid_constant__ DevCamera<float> cam, int width, int height,
                                                           const __grid_constant__ PhiloxKeys keys, unsigned long long n_rays,
                                                           unsigned long long *__restrict__ out) {
    using N = Num<float>;
    const SceneView<float> sc = view_of<float>(scene.base, scene);
    const float4 *filt = reinterpret_cast<const float4 *>(static_cast<const char *>(scene.base) + scene.filt_off);
    const int *far = reinterpret_cast<const int *>(static_cast<const char *>(scene.base) + scene.far_off);
    unsigned long long pairs = 0, epass = 0, fpass = 0, missed = 0, skipped = 0;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n_rays;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        Philox ph;
        ph.open(keys, (uint32_t)k, (uint32_t)(k >> 32), 0x7fffffffu);
        ph.block(0);
        const float u0 = N::uniform(ph.w[0], 0), u1 = N::uniform(ph.w[1], 0), u2 = N::uniform(ph.w[2], 0), u3 = N::uniform(ph.w[3], 0);
        ph.block(1);
        const float u4 = N::uniform(ph.w[0], 0), u5 = N::uniform(ph.w[1], 0), u6 = N::uniform(ph.w[2], 0), u7 = N::uniform(ph.w[3], 0);
        const int kind = (int)(k & 3ull);
        Vec3<float> o = cam.center, d;
        auto on_sphere = [&](int slot, float ua, float ub) {            // point on the surface of `slot`
            const float4 s = sc.geom[slot];
            const float z = 1.f - 2.f * ua, rr = sqrtf(fmaxf(0.f, 1.f - z * z)), phi = 6.2831853f * ub;
            Vec3<float> p;
            p.x = s.x + s.w * rr * cosf(phi); p.y = s.y + s.w * z; p.z = s.z + s.w * rr * sinf(phi);
            return p;
        };
        const int sa = min(scene.n - 1, (int)(u0 * scene.n)), sb = min(scene.n - 1, (int)(u1 * scene.n));
        if (kind == 0) {
            const float px = u0 * width, py = u1 * height;
            d.x = fmaf(py, cam.dv.x, fmaf(px, cam.du.x, cam.pixel00.x)) - o.x;
            d.y = fmaf(py, cam.dv.y, fmaf(px, cam.du.y, cam.pixel00.y)) - o.y;
            d.z = fmaf(py, cam.dv.z, fmaf(px, cam.du.z, cam.pixel00.z)) - o.z;
        } else if (kind == 1) {
            o = on_sphere(sa, u2, u3);
            const float z = 1.f - 2.f * u4, rr = sqrtf(fmaxf(0.f, 1.f - z * z)), phi = 6.2831853f * u5;
            const float len = 0.01f + 2.f * u6;
            d.x = len * rr * cosf(phi); d.y = len * z; d.z = len * rr * sinf(phi);
        } else {
            if (kind == 3) o = on_sphere(sa, u2, u3);
            // tangent point of sphere sb seen from o: c + r * (-(r/D) v + sqrt(1 - (r/D)^2) n), v = unit(c - o), n perp v
            const float4 s = sc.geom[sb];
            Vec3<float> v; v.x = s.x - o.x; v.y = s.y - o.y; v.z = s.z - o.z;
            const float D = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
            if (!(D > fabsf(s.w) * 1.0001f) || !(s.w > 0.f)) { ++skipped; continue; }
            v.x /= D; v.y /= D; v.z /= D;
            Vec3<float> e; e.x = u4 - .5f; e.y = u5 - .5f; e.z = u6 - .5f;           // random vector, made perpendicular to v
            const float ev = e.x * v.x + e.y * v.y + e.z * v.z;
            e.x -= ev * v.x; e.y -= ev * v.y; e.z -= ev * v.z;
            const float el = sqrtf(e.x * e.x + e.y * e.y + e.z * e.z);
            if (!(el > 1e-3f)) { ++skipped; continue; }
            e.x /= el; e.y /= el; e.z /= el;
            const float q = s.w / D, w = sqrtf(fmaxf(0.f, 1.f - q * q));
            const float nudge = 1.f + (u7 - .5f) * 4e-6f;                            // a few ulp either side of tangency
            Vec3<float> p;
            p.x = s.x + s.w * nudge * (-q * v.x + w * e.x); p.y = s.y + s.w * nudge * (-q * v.y + w * e.y); p.z = s.z + s.w * nudge * (-q * v.z + w * e.z);
            const float len = 0.05f + 3.f * u3 * u3;
            d.x = (p.x - o.x) * len; d.y = (p.y - o.y) * len; d.z = (p.z - o.z) * len;
        }
        const FilterRay fr = filter_ray(o, d, scene.bound);
        if (!fr.sane) { ++skipped; continue; }
        int next_far = 0;
        for (int i = 0; i < scene.n; ++i) {
            if (next_far < scene.n_far && far[next_far] == i) { ++next_far; continue; }   // far slots are not filtered
            const int half = i >= scene.n_half;
            const float4 q = filt[half * scene.half_pad + (i - half * scene.n_half)];
            float h;
            const float disc = disc_of<float>(sc.geom[i], o, d, fr.a, h);
            const bool e = disc >= 0.0f, f = filter_value(q, fr) >= fr.thr;
            ++pairs;
            epass += e; fpass += f; missed += (e && !f);
        }
    }
    unsigned long long v[5] = {pairs, epass, fpass, missed, skipped};
#pragma unroll
    for (int q = 0; q < 5; ++q) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[q] += __shfl_xor_sync(FULL, v[q], off);
        if ((threadIdx.x & 31) == 0 && v[q]) atomicAdd(out + q, v[q]);
    }
}

}  // namespace rt

// ==============================================================================================
// C ABI, device half
// ==============================================================================================
using namespace rt;

struct rt_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // scene
    void *scene_dev = nullptr;
    SceneBlob blob{};
    int scene_prec = 0;                 // 0 none, 4 float, 8 double
    // scratch: fixed-point accumulators of the last render, [pixels][3] int64
    void *acc = nullptr;
    size_t acc_bytes = 0;
    void *frame = nullptr;              // device frame when the caller's buffer is host memory
    size_t frame_bytes = 0;
    unsigned long long *queue = nullptr;   // 8 x u64: job cursor + work counters
    rt_stats stats{};
    // LBVH (float scenes; built on the device the first time RT_ACCEL_LBVH is asked for)
    std::vector<float4> host_geom;
    bool bvh_ready = false;
    BvhView bvh{};
    void *bvh_mem[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // nodes, geom, slot, big_geom, big_slot
    float bvh_build_ms = 0.f;
    // wavefront variant: path pool
    void *wf_mem = nullptr;
    size_t wf_bytes = 0;
    // uniform grid (RT_ACCEL_GRID, experimental): built on the host the first time it is asked for
    bool grid_ready = false, grid_usable = false, grid_auto = false;   // usable: the structure fits; auto: RT_ACCEL_AUTO may pick it
    float grid_build_ms = 0.f;
    GridView grid{};
    void *grid_mem[4] = {nullptr, nullptr, nullptr, nullptr};   // start, items, big_geom, big_slot
    // per-tile candidate lists of the camera rays (rt_primary_bins.cuh), rebuilt by every render call
    void *bins = nullptr;
    size_t bins_bytes = 0;
};

constexpr int QUEUE_WORDS = 8;

#define RT_CUDA(call)                                         \
    do {                                                      \
        cudaError_t err__ = (call);                           \
        if (err__ != cudaSuccess) return (int)err__;          \
    } while (0)

namespace {

bool is_device_ptr(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int ensure(void **buf, size_t *have, size_t need) {
    if (*have >= need) return 0;
    if (*buf) { cudaError_t e = cudaFree(*buf); *buf = nullptr; *have = 0; if (e != cudaSuccess) return (int)e; }
    cudaError_t e = cudaMalloc(buf, need);
    if (e != cudaSuccess) return (int)e;
    *have = need;
    return 0;
}

template <typename T, typename Cam> DevCamera<T> to_dev(const Cam &c) {
    DevCamera<T> d;
    d.center = {c.center[0], c.center[1], c.center[2]};
    d.pixel00 = {c.pixel00[0], c.pixel00[1], c.pixel00[2]};
    d.du = {c.du[0], c.du[1], c.du[2]};
    d.dv = {c.dv[0], c.dv[1], c.dv[2]};
    d.disk_u = {c.disk_u[0], c.disk_u[1], c.disk_u[2]};
    d.disk_v = {c.disk_v[0], c.disk_v[1], c.disk_v[2]};
    d.defocus_angle = c.defocus_angle;
    d.scale = c.scale;
    return d;
}

// Conservative filter data of the paired scan (rt_device.cuh).
//   far set   slots whose |centre| + radius is far above the rest (> 8 x the median, at most 16 of them) or not
//             finite: they would inflate the per-ray error bound of every slot, so they stay out of the filter
//             and are tested exactly for every ray (the 1000-unit ground sphere of the reference's scenes);
//   filt[]    (cx, cy, cz, -(|c|^2 - r^2)) for the other slots, the slot list cut into two halves (one per lane
//             of a pair), each padded to a multiple of 8 with records that never pass (nk = -inf);
//   bound     max |centre| + radius over the filtered slots, rounded up.
// The filter is switched off (exact scan) for scenes whose scale would let the bound over- or underflow.
struct FilterPlan {
    bool ok = false;
    int n_half = 0, half_pad = 0;
    float bound = 0.f;
    std::vector<float4> filt;
    std::vector<int> far;
};

FilterPlan plan_filter(const std::vector<float4> &g) {
    FilterPlan P;
    const int n = (int)g.size();
    if (getenv("RT_NO_FILTER")) return P;                                 // debugging aid: exact scan
    std::vector<double> ext((size_t)n);
    std::vector<double> finite_ext;
    for (int i = 0; i < n; ++i) {
        const float4 &s = g[(size_t)i];
        const double e = std::sqrt((double)s.x * s.x + (double)s.y * s.y + (double)s.z * s.z) + std::fabs((double)s.w);
        ext[(size_t)i] = e;
        if (std::isfinite(e)) finite_ext.push_back(e);
    }
    double median = 0.0;
    if (!finite_ext.empty()) {
        std::nth_element(finite_ext.begin(), finite_ext.begin() + (long)(finite_ext.size() / 2), finite_ext.end());
        median = finite_ext[finite_ext.size() / 2];
    }
    std::vector<char> is_far((size_t)n, 0);
    std::vector<std::pair<double, int>> big;
    for (int i = 0; i < n; ++i) {
        if (!std::isfinite(ext[(size_t)i])) { is_far[(size_t)i] = 1; P.far.push_back(i); }
        else if (ext[(size_t)i] > 8.0 * median) big.emplace_back(ext[(size_t)i], i);
    }
    std::sort(big.begin(), big.end(), [](const std::pair<double, int> &x, const std::pair<double, int> &y) { return x.first > y.first; });
    for (size_t k = 0; k < big.size() && P.far.size() < 16; ++k) { is_far[(size_t)big[k].second] = 1; P.far.push_back(big[k].second); }
    if (P.far.size() > 16) { P.far.clear(); return P; }                   // many non-finite slots: exact scan
    std::sort(P.far.begin(), P.far.end());
    double bound = 0.0;
    for (int i = 0; i < n; ++i) if (!is_far[(size_t)i]) bound = std::max(bound, ext[(size_t)i]);
    if (!(bound > 1e-12 && bound < 1e12)) {
        if (bound == 0.0 && (int)P.far.size() == n) bound = 1.0;          // nothing filtered
        else { P.far.clear(); return P; }
    }
    P.bound = (float)(bound * (1.0 + 1e-6));
    P.n_half = (n + 1) / 2;
    P.half_pad = (P.n_half + 7) & ~7;
    const float4 never = make_float4(0.f, 0.f, 0.f, -INFINITY);
    P.filt.assign((size_t)2 * P.half_pad, never);
    for (int i = 0; i < n; ++i) {
        if (is_far[(size_t)i]) continue;
        const float4 &s = g[(size_t)i];
        const double k = (double)s.x * s.x + (double)s.y * s.y + (double)s.z * s.z - (double)s.w * s.w;
        const int half = i >= P.n_half, idx = i - half * P.n_half;
        P.filt[(size_t)half * P.half_pad + idx] = make_float4(s.x, s.y, s.z, (float)(-k));
    }
    P.ok = true;
    return P;
}

template <typename T, typename Slot> int upload(rt_ctx *ctx, const Slot *slots, int n) {
    using V4 = typename Num<T>::vec4;
    if (!ctx || !slots || n <= 0 || n > (1 << 24)) return RT_EINVAL;
    RT_CUDA(cudaSetDevice(ctx->device));
    const size_t n8 = ((size_t)n + 7) & ~(size_t)7;                       // the scan's tail walks groups of 8 slots
    const size_t type_bytes = ((size_t)n * sizeof(int) + 15) & ~(size_t)15;
    const size_t geom_bytes = n8 * sizeof(V4);
    const size_t matl_bytes = (size_t)n * sizeof(V4);
    const size_t rinv_bytes = ((size_t)n * sizeof(T) + 15) & ~(size_t)15;
    // float scenes: the conservative filter of the paired scan (rt_device.cuh), built here once per scene
    FilterPlan plan;
    {
        std::vector<float4> g((size_t)n);
        for (int i = 0; i < n; ++i) g[(size_t)i] = make_float4((float)slots[i].cx, (float)slots[i].cy, (float)slots[i].cz, (float)slots[i].r);
        plan = plan_filter(g);
    }
    const size_t filt_bytes = plan.filt.size() * sizeof(float4);
    const size_t far_bytes = (plan.far.size() * sizeof(int) + 15) & ~(size_t)15;
    const size_t total = geom_bytes + matl_bytes + type_bytes + rinv_bytes + filt_bytes + far_bytes;
    std::vector<unsigned char> host(total, 0);
    V4 *geom = reinterpret_cast<V4 *>(host.data());
    V4 *matl = reinterpret_cast<V4 *>(host.data() + geom_bytes);
    int *type = reinterpret_cast<int *>(host.data() + geom_bytes + matl_bytes);
    T *rinv = reinterpret_cast<T *>(host.data() + geom_bytes + matl_bytes + type_bytes);
    for (int i = 0; i < n; ++i) {
        const Slot &s = slots[i];
        if (s.type < 0 || s.type > 2) return RT_EINVAL;
        geom[i].x = s.cx; geom[i].y = s.cy; geom[i].z = s.cz; geom[i].w = s.r;
        matl[i].x = s.albedo[0]; matl[i].y = s.albedo[1]; matl[i].z = s.albedo[2];
        matl[i].w = s.type == RT_METAL ? s.fuzz : (s.type == RT_DIELECTRIC ? s.ri : T(0));
        type[i] = s.type;
        rinv[i] = T(1) / static_cast<T>(s.r);                    // IEEE division == the device's rcp.rn
    }
    if (filt_bytes) std::memcpy(host.data() + geom_bytes + matl_bytes + type_bytes + rinv_bytes, plan.filt.data(), filt_bytes);
    if (!plan.far.empty())
        std::memcpy(host.data() + geom_bytes + matl_bytes + type_bytes + rinv_bytes + filt_bytes, plan.far.data(), plan.far.size() * sizeof(int));
    if (ctx->scene_dev) { RT_CUDA(cudaFree(ctx->scene_dev)); ctx->scene_dev = nullptr; }
    RT_CUDA(cudaMalloc(&ctx->scene_dev, total));
    RT_CUDA(cudaMemcpyAsync(ctx->scene_dev, host.data(), total, cudaMemcpyHostToDevice, ctx->stream));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->blob.base = ctx->scene_dev;
    ctx->blob.bytes = (uint32_t)total;
    ctx->blob.matl_off = (uint32_t)geom_bytes;
    ctx->blob.type_off = (uint32_t)(geom_bytes + matl_bytes);
    ctx->blob.rinv_off = (uint32_t)(geom_bytes + matl_bytes + type_bytes);
    ctx->blob.n = n;
    ctx->blob.filt_off = (uint32_t)(geom_bytes + matl_bytes + type_bytes + rinv_bytes);
    ctx->blob.far_off = (uint32_t)(geom_bytes + matl_bytes + type_bytes + rinv_bytes + filt_bytes);
    ctx->blob.filter_ok = plan.ok ? 1 : 0;
    ctx->blob.n_half = plan.n_half;
    ctx->blob.half_pad = plan.half_pad;
    ctx->blob.n_far = (int)plan.far.size();
    ctx->blob.bound = plan.bound;
    ctx->scene_prec = (int)sizeof(T);
    // a new scene invalidates the LBVH; keep the float geometry for its host-side classification
    for (void *&m : ctx->bvh_mem) if (m) { cudaFree(m); m = nullptr; }
    ctx->bvh_ready = false;
    ctx->bvh = BvhView{};
    for (void *&m : ctx->grid_mem) if (m) { cudaFree(m); m = nullptr; }
    ctx->grid_ready = ctx->grid_usable = ctx->grid_auto = false;
    ctx->grid = GridView{};
    // float copy of the geometry for the host-side builds (LBVH classification: float scenes only; uniform grid: double scenes
    // too -- the registration padding, 5 % of a cell, covers the rounding by orders of magnitude)
    ctx->host_geom.resize((size_t)n);
    for (int i = 0; i < n; ++i)
        ctx->host_geom[(size_t)i] = make_float4((float)slots[i].cx, (float)slots[i].cy, (float)slots[i].cz, (float)slots[i].r);
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// LBVH build: classification on the host (which spheres are too big for the tree), everything
// else -- Morton codes, radix sort, hierarchy, refit -- on the device.
int build_lbvh(rt_ctx *ctx) {
    if (ctx->bvh_ready) return RT_OK;
    if (ctx->scene_prec != 4 || ctx->host_geom.empty()) return RT_EPRECISION;
    const int n = ctx->blob.n;
    const std::vector<float4> &g = ctx->host_geom;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (const float4 &s : g) {
        lo[0] = std::min(lo[0], s.x); lo[1] = std::min(lo[1], s.y); lo[2] = std::min(lo[2], s.z);
        hi[0] = std::max(hi[0], s.x); hi[1] = std::max(hi[1], s.y); hi[2] = std::max(hi[2], s.z);
    }
    const float diag = std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) +
                                 (hi[2] - lo[2]) * (hi[2] - lo[2]));
    // spheres far larger than the scene (the ground) and far smaller than the rest (the reference's never-written slot is a
    // zero-radius sphere; it would force the smallest-radius inflation of every box above it) stay out of the tree
    std::vector<float> radii((size_t)n);
    for (int i = 0; i < n; ++i) radii[(size_t)i] = std::fabs(g[(size_t)i].w);
    std::nth_element(radii.begin(), radii.begin() + n / 2, radii.end());
    const float r_med = radii[(size_t)(n / 2)];
    std::vector<int> small_idx, big_idx;
    for (int i = 0; i < n; ++i) {
        const float r = std::fabs(g[(size_t)i].w);
        ((r > 0.25f * diag || r < 0.05f * r_med) ? big_idx : small_idx).push_back(i);
    }
    if (big_idx.size() > 64) { small_idx.resize((size_t)n); for (int i = 0; i < n; ++i) small_idx[(size_t)i] = i; big_idx.clear(); }
    const int m = (int)small_idx.size(), nbig = (int)big_idx.size();
    for (int q = 0; q < 3; ++q) { lo[q] = INFINITY; hi[q] = -INFINITY; }
    for (int i : small_idx) {
        const float4 &s = g[(size_t)i];
        lo[0] = std::min(lo[0], s.x); lo[1] = std::min(lo[1], s.y); lo[2] = std::min(lo[2], s.z);
        hi[0] = std::max(hi[0], s.x); hi[1] = std::max(hi[1], s.y); hi[2] = std::max(hi[2], s.z);
    }
    const float4 *geom_dev = static_cast<const float4 *>(ctx->scene_dev);
    cudaStream_t st = ctx->stream;
    RT_CUDA(cudaEventRecord(ctx->ev[3], st));

    // persistent arrays; `keep` frees them on every failing exit path, `tmp` releases the device temporaries on every path
    float4 *nodes = nullptr, *geom_sorted = nullptr, *big_geom = nullptr;
    int *slot_sorted = nullptr, *big_slot = nullptr;
    struct Keep {                       // the arrays that outlive the build; freed here only if the build fails
        void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        bool armed = true;
        ~Keep() { if (armed) for (void *q : p) if (q) cudaFree(q); }
    } keep;
    struct Scratch {
        std::vector<void *> ptrs;
        ~Scratch() { for (void *p : ptrs) cudaFree(p); }
        cudaError_t get(void **p, size_t bytes) {
            const cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
            if (e == cudaSuccess) ptrs.push_back(*p);
            return e;
        }
    } tmp;
    if (nbig) {
        std::vector<float4> bg((size_t)nbig);
        for (int b = 0; b < nbig; ++b) bg[(size_t)b] = g[(size_t)big_idx[(size_t)b]];
        RT_CUDA(cudaMalloc(&big_geom, sizeof(float4) * nbig));
        keep.p[3] = big_geom;
        RT_CUDA(cudaMalloc(&big_slot, sizeof(int) * nbig));
        keep.p[4] = big_slot;
        RT_CUDA(cudaMemcpyAsync(big_geom, bg.data(), sizeof(float4) * nbig, cudaMemcpyHostToDevice, st));
        RT_CUDA(cudaMemcpyAsync(big_slot, big_idx.data(), sizeof(int) * nbig, cudaMemcpyHostToDevice, st));
        RT_CUDA(cudaStreamSynchronize(st));
    }
    if (m > 0) {
        RT_CUDA(cudaMalloc(&geom_sorted, sizeof(float4) * m));
        keep.p[1] = geom_sorted;
        RT_CUDA(cudaMalloc(&slot_sorted, sizeof(int) * m));
        keep.p[2] = slot_sorted;
        RT_CUDA(cudaMalloc(&nodes, sizeof(float4) * 4 * (size_t)std::max(1, m - 1)));
        keep.p[0] = nodes;
        int *small_dev = nullptr, *vals = nullptr, *parent = nullptr, *flags = nullptr;
        uint32_t *keys = nullptr, *keys_sorted = nullptr;
        float *box = nullptr, *rad = nullptr;
        int2 *children = nullptr;
        void *cub_tmp = nullptr;
        size_t cub_bytes = 0;
        const size_t nn = (size_t)2 * m - 1;
        RT_CUDA(tmp.get((void **)&small_dev, sizeof(int) * m));
        RT_CUDA(tmp.get((void **)&vals, sizeof(int) * m));
        RT_CUDA(tmp.get((void **)&keys, sizeof(uint32_t) * m));
        RT_CUDA(tmp.get((void **)&keys_sorted, sizeof(uint32_t) * m));
        RT_CUDA(tmp.get((void **)&box, sizeof(float) * 6 * nn));
        RT_CUDA(tmp.get((void **)&rad, sizeof(float) * nn));
        RT_CUDA(tmp.get((void **)&parent, sizeof(int) * nn));
        RT_CUDA(tmp.get((void **)&flags, sizeof(int) * std::max(1, m - 1)));
        RT_CUDA(tmp.get((void **)&children, sizeof(int2) * std::max(1, m - 1)));
        RT_CUDA(cudaMemcpyAsync(small_dev, small_idx.data(), sizeof(int) * m, cudaMemcpyHostToDevice, st));
        const int tb = 256, gb = (m + tb - 1) / tb;
        const float3 lo3 = make_float3(lo[0], lo[1], lo[2]);
        const float3 inv3 = make_float3(hi[0] > lo[0] ? 1.f / (hi[0] - lo[0]) : 0.f, hi[1] > lo[1] ? 1.f / (hi[1] - lo[1]) : 0.f,
                                        hi[2] > lo[2] ? 1.f / (hi[2] - lo[2]) : 0.f);
        bvh_morton_kernel<<<gb, tb, 0, st>>>(geom_dev, small_dev, m, lo3, inv3, keys, vals);
        RT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys, keys_sorted, vals, slot_sorted, m, 0, 30, st));
        RT_CUDA(tmp.get(&cub_tmp, cub_bytes));
        RT_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys, keys_sorted, vals, slot_sorted, m, 0, 30, st));
        bvh_leaves_kernel<<<gb, tb, 0, st>>>(geom_dev, slot_sorted, m, geom_sorted, box, rad);
        if (m > 1) {
            RT_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * (m - 1), st));
            bvh_hierarchy_kernel<<<(m - 1 + tb - 1) / tb, tb, 0, st>>>(keys_sorted, m, children, parent);
            bvh_refit_kernel<<<gb, tb, 0, st>>>(m, children, parent, box, rad, flags, nodes);
        }
        RT_CUDA(cudaGetLastError());
        RT_CUDA(cudaStreamSynchronize(st));
    }
    RT_CUDA(cudaEventRecord(ctx->ev[2], st));
    RT_CUDA(cudaEventSynchronize(ctx->ev[2]));
    RT_CUDA(cudaEventElapsedTime(&ctx->bvh_build_ms, ctx->ev[3], ctx->ev[2]));
    keep.armed = false;
    ctx->bvh_mem[0] = nodes; ctx->bvh_mem[1] = geom_sorted; ctx->bvh_mem[2] = slot_sorted;
    ctx->bvh_mem[3] = big_geom; ctx->bvh_mem[4] = big_slot;
    ctx->bvh.nodes = nodes; ctx->bvh.geom = geom_sorted; ctx->bvh.slot = slot_sorted;
    ctx->bvh.big_geom = big_geom; ctx->bvh.big_slot = big_slot;
    ctx->bvh.m = m; ctx->bvh.nbig = nbig;
    // bounds and smallest radius of the tree's spheres.  The scene counts as compact when the single per-ray box inflation
    // (bvh_start<true>) stays below 10 % of the smallest radius for every origin within one bounds-diagonal of the tree.
    float rmin_all = INFINITY, diag2 = 0.f;
    for (int q = 0; q < 3; ++q) { ctx->bvh.blo[q] = INFINITY; ctx->bvh.bhi[q] = -INFINITY; }
    for (int i : small_idx) {
        const float4 &sph = g[(size_t)i];
        const float c[3] = {sph.x, sph.y, sph.z}, rr = std::fabs(sph.w);
        for (int q = 0; q < 3; ++q) { ctx->bvh.blo[q] = std::min(ctx->bvh.blo[q], c[q] - rr); ctx->bvh.bhi[q] = std::max(ctx->bvh.bhi[q], c[q] + rr); }
        rmin_all = std::min(rmin_all, rr);
    }
    for (int q = 0; q < 3; ++q) diag2 += (ctx->bvh.bhi[q] - ctx->bvh.blo[q]) * (ctx->bvh.bhi[q] - ctx->bvh.blo[q]);
    ctx->bvh.rmin_all = (m > 0 && std::isfinite(rmin_all)) ? rmin_all : 0.f;
    ctx->bvh.compact = 0;
    if (m > 1 && ctx->bvh.rmin_all > 0.f && std::isfinite(diag2)) {
        const float reach = 2.f * std::sqrt(diag2);
        const float delta = std::sqrt(BVH_KEPS * reach * reach + rmin_all * rmin_all) - rmin_all;
        ctx->bvh.compact = (delta <= 0.10f * rmin_all && !getenv("RT_BVH_NO_COMPACT")) ? 1 : 0;
    }
    ctx->bvh_ready = true;
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// Uniform grid (rt_grid.cuh), host build; mirrors tools/grid_model.py (class Grid).  Sets ctx->grid_usable = false for scenes
// the structure does not fit (fewer than two similar spheres, more than 64 spheres of a very different size).
int build_grid(rt_ctx *ctx) {
    if (ctx->grid_ready) return RT_OK;
    if (ctx->host_geom.empty()) return RT_ENOSCENE;
    const std::vector<float4> &g = ctx->host_geom;
    const int n = (int)g.size();
    std::vector<double> rad;
    for (const float4 &s : g) {
        const double r = std::fabs((double)s.w);
        if (std::isfinite(s.x) && std::isfinite(s.y) && std::isfinite(s.z) && std::isfinite(r) && r > 0.0) rad.push_back(r);
    }
    double r_med = 0.0;
    if (!rad.empty()) { std::sort(rad.begin(), rad.end()); r_med = rad.size() % 2 ? rad[rad.size() / 2] : 0.5 * (rad[rad.size() / 2 - 1] + rad[rad.size() / 2]); }
    // Field spheres (radius within a factor 4 of the median) go into the grid; everything else -- the ground, the reference's
    // three unit spheres, the never-written zero-radius slot, non-finite records -- is tested for every ray.  (Registering the
    // unit spheres in the cells they overlap was tried: the slab the walk is clipped to grows from 0.4 to 2 units, and the
    // CPU model visits 2-4x the cells per segment -- more than the three exact tests it saves.)
    std::vector<int> in_grid, big;
    for (int i = 0; i < n; ++i) {
        const float4 &s = g[(size_t)i];
        const double r = std::fabs((double)s.w);
        const bool fin = std::isfinite(s.x) && std::isfinite(s.y) && std::isfinite(s.z) && std::isfinite(r);
        if (fin && r > 0.0 && r >= 0.25 * r_med && r <= 4.0 * r_med) in_grid.push_back(i);
        else big.push_back(i);
    }
    if (in_grid.size() < 2 || big.size() > 64) {                      // not a field of similar spheres
        ctx->grid_ready = true;
        ctx->grid_usable = ctx->grid_auto = false;
        return RT_OK;
    }
    const auto t_build0 = std::chrono::steady_clock::now();
    double cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    double lo3[3] = {INFINITY, INFINITY, INFINITY}, hi3[3] = {-INFINITY, -INFINITY, -INFINITY};
    double r_max = 0.0, r_min = INFINITY;
    auto grow = [&](int i) {
        const float4 &s = g[(size_t)i];
        const double c[3] = {s.x, s.y, s.z}, r = std::fabs((double)s.w);
        for (int q = 0; q < 3; ++q) { lo3[q] = std::min(lo3[q], c[q] - r); hi3[q] = std::max(hi3[q], c[q] + r); }
    };
    for (int i : in_grid) {
        const float4 &s = g[(size_t)i];
        const double c[3] = {s.x, s.y, s.z}, r = std::fabs((double)s.w);
        for (int q = 0; q < 3; ++q) { cmin[q] = std::min(cmin[q], c[q]); cmax[q] = std::max(cmax[q], c[q]); }
        grow(i);
        r_max = std::max(r_max, r); r_min = std::min(r_min, r);
    }
    int av = 0;
    for (int q = 1; q < 3; ++q) if (cmax[q] - cmin[q] < cmax[av] - cmin[av]) av = q;
    const int au = av == 0 ? 1 : 0, aw = av == 2 ? 1 : 2;
    double h = std::sqrt(std::max((hi3[au] - lo3[au]) * (hi3[aw] - lo3[aw]), 1e-30) / (double)in_grid.size());
    h = std::max(h, r_max);
    struct Undo {                                                     // a failed upload leaves no half-built grid behind
        rt_ctx *c; bool armed = true;
        ~Undo() { if (armed) { for (void *&m : c->grid_mem) if (m) { cudaFree(m); m = nullptr; } c->grid = GridView{}; } }
    } undo{ctx};
    const double eu = hi3[au] - lo3[au], ew = hi3[aw] - lo3[aw];
    const int nu = (int)std::min(std::max(std::ceil(eu / h), 1.0), 4096.0), nw = (int)std::min(std::max(std::ceil(ew / h), 1.0), 4096.0);
    h = std::max(std::max(eu / nu, ew / nw), h);
    GridView G{};
    G.nu = nu; G.nw = nw; G.au = au; G.av = av; G.aw = aw;
    G.h = (float)(h * (1.0 + 1e-6));
    G.inv_h = 1.0f / G.h;
    for (int q = 0; q < 3; ++q) { G.lo[q] = std::nextafter((float)lo3[q], -INFINITY); G.hi[q] = std::nextafter((float)hi3[q], INFINITY); }
    G.rmin = (float)r_min;
    G.pad = (float)(0.05 * (double)G.h);
    G.half_pad = G.pad * 0.5f;
    G.ulo = G.lo[au];
    G.wlo = G.lo[aw];
    const double hh = G.h, ulo = G.lo[au], wlo = G.lo[aw];
    std::vector<unsigned int> count((size_t)nu * nw + 1, 0u);
    auto range = [&](double c, double R, double base, int cells, int &a0, int &a1) {
        a0 = (int)std::min(std::max(std::floor((c - R - base) / hh), 0.0), (double)(cells - 1));
        a1 = (int)std::min(std::max(std::floor((c + R - base) / hh), 0.0), (double)(cells - 1));
    };
    std::vector<unsigned int> items, fill;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            unsigned int run = 0;
            for (size_t k = 0; k < count.size(); ++k) { const unsigned int c = count[k]; count[k] = run; run += c; }   // exclusive scan
            items.assign(run, 0u);
            fill.assign(count.begin(), count.end());
        }
        for (int i : in_grid) {
            const float4 &s = g[(size_t)i];
            const double c[3] = {s.x, s.y, s.z};
            const double R = (std::fabs((double)s.w) + (double)G.pad) * (1.0 + 1e-6);
            int u0, u1, w0, w1;
            range(c[au], R, ulo, nu, u0, u1);
            range(c[aw], R, wlo, nw, w0, w1);
            for (int b = w0; b <= w1; ++b)
                for (int a = u0; a <= u1; ++a) {
                    const size_t cell = (size_t)b * nu + a;
                    if (pass == 0) ++count[cell];
                    else items[fill[cell]++] = (unsigned int)i;
                }
        }
        if (pass == 1) {
            std::vector<float4> bg(big.size());
            for (size_t b = 0; b < big.size(); ++b) bg[b] = g[(size_t)big[b]];
            RT_CUDA(cudaMalloc(&ctx->grid_mem[0], count.size() * sizeof(unsigned int)));
            RT_CUDA(cudaMalloc(&ctx->grid_mem[1], std::max<size_t>(items.size(), 1) * sizeof(unsigned int)));
            RT_CUDA(cudaMalloc(&ctx->grid_mem[2], std::max<size_t>(bg.size(), 1) * sizeof(float4)));
            RT_CUDA(cudaMalloc(&ctx->grid_mem[3], std::max<size_t>(big.size(), 1) * sizeof(int)));
            RT_CUDA(cudaMemcpyAsync(ctx->grid_mem[0], count.data(), count.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream));
            if (!items.empty()) RT_CUDA(cudaMemcpyAsync(ctx->grid_mem[1], items.data(), items.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream));
            if (!bg.empty()) RT_CUDA(cudaMemcpyAsync(ctx->grid_mem[2], bg.data(), bg.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
            if (!big.empty()) RT_CUDA(cudaMemcpyAsync(ctx->grid_mem[3], big.data(), big.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            RT_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    G.start = static_cast<const unsigned int *>(ctx->grid_mem[0]);
    G.items = static_cast<const unsigned int *>(ctx->grid_mem[1]);
    G.big_geom = static_cast<const float4 *>(ctx->grid_mem[2]);
    G.big_slot = static_cast<const int *>(ctx->grid_mem[3]);
    G.nbig = (int)big.size();
    G.n_items = (unsigned int)items.size();
    G.n_slots = n;
    ctx->grid = G;
    undo.armed = false;
    ctx->grid_ready = true;
    ctx->grid_usable = true;
    // RT_ACCEL_AUTO picks the grid for a planar field of moderate extent: the slab axis is thin against the cell size, and the
    // float-noise inflation delta of a sphere half a diagonal away from the ray origin stays within half a cell, so that a step
    // looks at its own cell and -- far from the origin only -- at one ring of neighbours.  Measured at 1920x1080
    // (tools/time_accels.py, tools/ab_variants.sh; grid / LBVH / scan): 40 slots 17.7 / 19.2 / 21.2 ms, 125 slots 20.1 / 23.9 /
    // 29.4, 488 slots 34.5 / 48.3 / 79.8, 3 604 slots 45 / 72 / 630, 14 404 slots 35.6 / 57.5, 99 860 slots (far cells need the
    // ring) 39.7 / 41.9, and at 3840x2160 / 128 spp 565 / 620 ms (profiles/logs/r02ag_cfg5_modes.log, r02ah_coop_ab.log).  Until
    // the last measurement the 99 860-slot field was left to the LBVH (delta within the registration padding was the rule); larger
    // fields than that have not been measured and still are.
    double diag2 = 0.0;
    for (int q = 0; q < 3; ++q) diag2 += (hi3[q] - lo3[q]) * (hi3[q] - lo3[q]);
    const double reach = 0.5 * std::sqrt(diag2);
    const double delta_half = (std::sqrt((double)BVH_KEPS * reach * reach + r_min * r_min) - r_min) * 1.001 + 1e-7 + 4.8e-7 * 2.0 * reach;
    ctx->grid_auto = (hi3[av] - lo3[av]) <= 4.0 * (double)G.h && delta_half <= 0.5 * (double)G.h && std::isfinite(delta_half);
    ctx->grid_build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_build0).count();
    return RT_OK;
}

// RT_ACCEL_AUTO -> the fastest structure for this scene and these options (same image, bit for bit, whichever is chosen).
// Float scenes from RT_AUTO_MIN_SLOTS slots up go through the uniform grid when they are a compact field of similar spheres
// (build_grid: grid_auto) and through the LBVH otherwise; small scenes, double scenes and the wavefront variant keep the
// shared-memory scan.
constexpr int GRID_BINS_FROM_BVH = 8192;    // above this the grid's tile lists come from a walk of the LBVH (float scenes only)
constexpr int RT_AUTO_MIN_SLOTS = 32;     // measured (tools/time_accels.py): the grid wins from 40 slots up (scene 2: 19.2 / 20.2 / 21.4 ms grid / LBVH / scan)
int resolve_accel(rt_ctx *ctx, const rt_opts &o) {
    if (o.accel != RT_ACCEL_AUTO) return o.accel;
    if (o.kernel != RT_KERNEL_MEGA) return RT_ACCEL_LINEAR;
    int min_slots = RT_AUTO_MIN_SLOTS;
    if (const char *e = getenv("RT_AUTO_MIN_SLOTS")) min_slots = atoi(e);                    // tuning knob
    if (ctx->blob.n < min_slots) return RT_ACCEL_LINEAR;
    const bool is_double = ctx->scene_prec != 4;
    if (o.primary_bins != RT_PBINS_OFF && !getenv("RT_AUTO_NO_GRID") && !(is_double && ctx->blob.n > GRID_BINS_FROM_BVH) &&
        build_grid(ctx) == RT_OK && ctx->grid_usable && ctx->grid_auto)
        return RT_ACCEL_GRID;
    return is_double ? RT_ACCEL_LINEAR : RT_ACCEL_LBVH;                 // the LBVH is a float structure
}
int resolve_accel(rt_ctx *ctx, int accel) {
    rt_opts o;
    rt_opts_default(&o);
    o.accel = accel;
    return resolve_accel(ctx, o);
}

// The LBVH is a float structure: these helpers keep the double instantiation of trace() from naming float-only kernels
// (trace() rejects double + LBVH before it gets here).
inline void launch_bins_bvh(const DevCamera<float> &cam, const BvhView &bv, int w, int h, int tx, int ty, unsigned int *bins, unsigned grid,
                            cudaStream_t st) {
    bin_kernel_bvh<<<grid, 128, 0, st>>>(cam, bv, w, h, tx, ty, bins);
}
inline void launch_bins_bvh(const DevCamera<double> &, const BvhView &, int, int, int, int, unsigned int *, unsigned, cudaStream_t) {}

size_t trace_smem(const SceneBlob &b) { return (size_t)b.bytes + (size_t)CAND_CAP * TRACE_BLOCK * sizeof(unsigned short); }

template <typename Kernel> int launch_shape(rt_ctx *ctx, Kernel kernel, size_t smem, int *grid) {
    RT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    RT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TRACE_BLOCK, smem));
    if (per_sm < 1) return RT_EINVAL;
    *grid = ctx->sm_count * per_sm;
    cudaFuncAttributes fa;
    RT_CUDA(cudaFuncGetAttributes(&fa, kernel));
    ctx->stats.regs = fa.numRegs;
    ctx->stats.smem_bytes = (int)(smem + fa.sharedSizeBytes);
    ctx->stats.grid = *grid;
    ctx->stats.block = TRACE_BLOCK;
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// Job plan (JobPlan, decode_job): how the samples [s_begin, s_begin + s_count) of `rows_local` x `width` pixels are cut into
// jobs.  Scheduling only -- the image does not depend on it.
int plan_jobs(int width, int rows_local, int s_begin, int s_count, JobPlan *out) {
    JobPlan P{};
    P.pix_local = (unsigned long long)rows_local * (unsigned long long)width;
    P.s_begin = s_begin;
    P.s_count = s_count;
    if (P.pix_local == 0 || s_count <= 0) { P.total_jobs = 0; *out = P; return RT_OK; }
    if (P.pix_local > 0xffffffffull) return RT_EINVAL;                     // pixel indices are 32-bit (Philox counter word)
    if ((long long)s_begin + s_count > (1ll << 30)) return RT_EINVAL;      // sample indices stay clear of int overflow
    int c_a = rt_num_chunks(width, rows_local, s_count);
    if (const char *e = getenv("RT_CHUNKS")) { const int v = atoi(e); if (v > 0) c_a = v < s_count ? v : s_count; }   // tuning knob
    // bands: as many rows as keep the band's accumulators (24 B per pixel) within 8 MB of L2
    long long band_rows = (8ll << 20) / (24ll * width);
    if (const char *e = getenv("RT_BAND_ROWS")) { const int v = atoi(e); if (v > 0) band_rows = v; }
    if (band_rows < 1) band_rows = 1;
    if (band_rows > rows_local) band_rows = rows_local;
    // 8 x 4 tile numbering needs whole tiles: bands of a multiple of 4 rows
    const bool tiled = width % 8 == 0 && rows_local % 4 == 0 && !getenv("RT_NO_TILE_ORDER");
    if (tiled) { band_rows = band_rows >= 4 ? band_rows / 4 * 4 : 4; }
    P.tiles_per_row = tiled ? (unsigned int)(width / 8) : 0u;
    const unsigned long long band_pix = (unsigned long long)band_rows * width;
    const unsigned long long bands = (P.pix_local + band_pix - 1) / band_pix;
    P.band_pix = (unsigned int)band_pix;
    P.last_pix = (unsigned int)(P.pix_local - (bands - 1) * band_pix);
    P.rp_b = (bands - 1) * band_pix;
    for (;; c_a = (c_a + 1) / 2) {
        P.spj = (s_count + c_a - 1) / c_a;
        P.chunks = (s_count + P.spj - 1) / P.spj;
        P.band_jobs = band_pix * (unsigned long long)P.chunks;
        P.jobs_a = (bands - 1) * P.band_jobs;
        P.total_jobs = P.jobs_a + (unsigned long long)P.last_pix * (unsigned long long)P.chunks;
        // multiply-high division is exact while dividend * divisor < 2^64; keep a wide margin.  Gigantic renders get longer jobs.
        if ((long double)P.total_jobs * (long double)P.band_jobs < 9.0e18L) break;
        if (c_a == 1) return RT_EINVAL;
    }
    auto magic = [](unsigned long long d) { return d > 1 ? ~0ull / d + 1ull : 0ull; };
    P.magic_band_jobs = magic(P.band_jobs);
    P.magic_pix[0] = magic(band_pix);
    P.magic_pix[1] = magic(P.last_pix);
    P.magic_width = magic((unsigned long long)width);
    P.magic_tpr = magic((unsigned long long)P.tiles_per_row);
    *out = P;
    return RT_OK;
}

// Fills the launch arguments every path-tracing kernel family shares and resets the queue / work counters.
template <typename T, typename Cam>
int fill_args(rt_ctx *ctx, const Cam &cam, const rt_opts &o, int rows_local, int s_begin, int s_count, long long *acc, TraceArgs<T> &A) {
    A.bvh = BvhView{};
    A.bvh_steps = 0;
    A.bvh_min_active = 0;
    A.bins = nullptr;
    A.tiles_x = 0;
    // camera-ray rounds per loop turn.  The shared-memory scan is expensive and must run with full warps: 3 rounds (measured at
    // config 2: 1 round 89-91 ms, 2 rounds 87.5-88, 3+ rounds 86.5-87); the grid walk and the LBVH traversal are cheap next to
    // a round: 1 round (36.69 / 36.85 ms grid, 50.01 / 50.22 ms LBVH for 1 / 3 rounds, profiles/logs/r02l_tile_capacity_rounds.log)
    A.pb_rounds = resolve_accel(ctx, o) == RT_ACCEL_LINEAR ? 3 : 1;
    A.pb_min = 1;
    if (const char *e = getenv("RT_PB_ROUNDS")) A.pb_rounds = atoi(e) > 0 ? atoi(e) : A.pb_rounds;       // tuning knobs
    if (const char *e = getenv("RT_PB_MIN")) A.pb_min = atoi(e);
    // lanes out of work claim together (adjacent pixels): measured 1 / 4 / 8 / 12 / 16 / 24 lanes -- grid 36.21 / 35.96 / 35.63 /
    // 35.66 / 36.21 / 40.2 ms, LBVH (99 860 slots) 43.02 / 42.79 / 42.93 / 43.99 / 46.35 / 58.25, scan 80.06 / 79.94 / 80.43 /
    // 82.12 / 85.46 / 103.5 (profiles/logs/r02z_cohort.log): waiting costs the scan more than coherence gives it
    {
        const int acc = resolve_accel(ctx, o);
        A.pb_cohort = acc == RT_ACCEL_GRID ? 8 : (acc == RT_ACCEL_LBVH ? 4 : 1);
    }
    if (const char *e = getenv("RT_PB_COHORT")) A.pb_cohort = atoi(e);
    A.cam = to_dev<T>(cam);
    A.scene = ctx->blob;
    A.keys = philox_keys(o.seed);
    A.max_depth = cam.max_depth;
    A.width = cam.width;
    A.height = cam.height;
    A.tiles = 0;
    A.tile_rows = o.tile_rows; A.rank = o.rank;
    A.world = (o.split == RT_SPLIT_ROWS) ? o.world : 1;
    const int rc = plan_jobs(cam.width, rows_local, s_begin, s_count, &A.plan);
    if (rc) return rc;
    A.acc = acc;
    A.queue = ctx->queue;
    RT_CUDA(cudaMemsetAsync(ctx->queue, 0, QUEUE_WORDS * sizeof(unsigned long long), ctx->stream));
    ctx->stats.chunks = A.plan.chunks;
    return RT_OK;
}

// camera-ray candidate lists of this frame (rt_primary_bins.cuh) into ctx->bins
template <typename T>
int build_bins(rt_ctx *ctx, TraceArgs<T> &A, int width, int height, bool from_bvh) {
    const int tiles_x = (width + (1 << PB_SHIFT) - 1) >> PB_SHIFT, tiles_y = (height + (1 << PB_SHIFT) - 1) >> PB_SHIFT;
    const size_t tiles = (size_t)tiles_x * tiles_y;
    const int rc = ensure(&ctx->bins, &ctx->bins_bytes, tiles * PB_STRIDE * sizeof(unsigned int));
    if (rc) return rc;
    unsigned int *bins = static_cast<unsigned int *>(ctx->bins);
    A.bins = bins;
    A.tiles_x = tiles_x;
    A.tiles = (int)tiles;
    const unsigned bin_grid = (unsigned)((tiles + 127) / 128);
    if (from_bvh) launch_bins_bvh(A.cam, ctx->bvh, width, height, tiles_x, tiles_y, bins, bin_grid, ctx->stream);
    else bin_kernel<T><<<bin_grid, 128, 0, ctx->stream>>>(A.cam, static_cast<const typename Num<T>::vec4 *>(ctx->blob.base), ctx->blob.n,
                                                         width, height, tiles_x, tiles_y, bins);
    RT_CUDA(cudaGetLastError());
    ctx->stats.launches += 1;
    return RT_OK;
}

// Wavefront variant: same jobs, same accumulators, two kernels per loop turn over a global pool.
template <typename T, typename Cam> struct WavefrontImpl {
    static int run(rt_ctx *, const Cam &, const rt_opts &, int, int, int, long long *) { return RT_EPRECISION; }
};
template <typename Cam> struct WavefrontImpl<float, Cam> {
    static int run(rt_ctx *ctx, const Cam &cam, const rt_opts &o, int rows_local, int s_begin, int s_count, long long *acc) {
        const size_t smem_hit = trace_smem(ctx->blob), smem_shade = ctx->blob.bytes;
        if (smem_hit > 227 * 1024 || ctx->blob.n > 65535) return RT_EINVAL;
        TraceArgs<float> A;
        int rc = fill_args<float>(ctx, cam, o, rows_local, s_begin, s_count, acc, A);
        if (rc) return rc;
        if (A.plan.total_jobs == 0) return RT_OK;
        // pool: 8 slots per resident thread of the megakernel's shape keeps the state near L2 size
        const unsigned long long want = (unsigned long long)ctx->sm_count * 2048ull * 4ull;
        const int n = (int)std::min<unsigned long long>(want, (A.plan.total_jobs + 255ull) & ~255ull);
        const size_t words = 18;                                   // 4-byte arrays per slot
        const size_t bytes = (size_t)n * 4 * (words + 2 * WF_CLASSES) + 512;
        rc = ensure(&ctx->wf_mem, &ctx->wf_bytes, bytes);
        if (rc) return rc;
        char *base = static_cast<char *>(ctx->wf_mem);
        auto take = [&](size_t elems, size_t elem_size) { void *p = base; base += ((elems * elem_size + 15) & ~(size_t)15); return p; };
        WfPool P;
        P.n = n;
        float **fl[] = {&P.ox, &P.oy, &P.oz, &P.dx, &P.dy, &P.dz, &P.ax, &P.ay, &P.az, &P.puy, &P.hit_t};
        for (float **f : fl) *f = static_cast<float *>(take((size_t)n, 4));
        int **in[] = {&P.hit_id, &P.sample, &P.sample_end, &P.depth, &P.state};
        for (int **f : in) *f = static_cast<int *>(take((size_t)n, 4));
        P.pixel = static_cast<uint32_t *>(take((size_t)n, 4));
        P.local = static_cast<uint32_t *>(take((size_t)n, 4));
        int *lists[2] = {static_cast<int *>(take((size_t)n * WF_CLASSES, 4)), static_cast<int *>(take((size_t)n * WF_CLASSES, 4))};
        unsigned int *counts = static_cast<unsigned int *>(take(16, 4));     // [2][4] counts, then alive
        unsigned int *cnt[2] = {counts, counts + WF_CLASSES};
        P.alive = counts + 2 * WF_CLASSES;
        P.list_in = lists[0]; P.list_out = lists[1];
        P.count_in = cnt[0]; P.count_out = cnt[1];
        RT_CUDA(cudaFuncSetAttribute(wf_shade, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_shade));
        RT_CUDA(cudaFuncSetAttribute(wf_intersect, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_hit));
        const int gb = (n + 255) / 256, gs = (n + 32 * WF_CLASSES + 255) / 256;
        wf_init<<<gb, 256, 0, ctx->stream>>>(P);
        cudaFuncAttributes fa;
        RT_CUDA(cudaFuncGetAttributes(&fa, wf_intersect));
        ctx->stats.regs = fa.numRegs; ctx->stats.smem_bytes = (int)smem_hit; ctx->stats.grid = gb; ctx->stats.block = 256;
        ctx->stats.launches += 1;
        int cur = 0;
        for (;;) {
            const int batch = 16;
            for (int it = 0; it < batch; ++it) {
                P.list_in = lists[cur]; P.list_out = lists[cur ^ 1];
                P.count_in = cnt[cur]; P.count_out = cnt[cur ^ 1];
                wf_shade<<<gs, 256, smem_shade, ctx->stream>>>(A, P);
                RT_CUDA(cudaMemsetAsync(P.count_out, 0, WF_CLASSES * sizeof(unsigned int), ctx->stream));
                RT_CUDA(cudaMemsetAsync(P.alive, 0, sizeof(unsigned int), ctx->stream));
                wf_intersect<<<gb, 256, smem_hit, ctx->stream>>>(A, P);
                cur ^= 1;
                ctx->stats.launches += 2;
            }
            RT_CUDA(cudaGetLastError());
            unsigned int alive = 0;
            RT_CUDA(cudaMemcpyAsync(&alive, P.alive, sizeof alive, cudaMemcpyDeviceToHost, ctx->stream));
            RT_CUDA(cudaStreamSynchronize(ctx->stream));
            if (alive == 0) break;
        }
        return RT_OK;
    }
};
template <typename T, typename Cam>
int trace_wavefront(rt_ctx *ctx, const Cam &cam, const rt_opts &o, int rows_local, int s_begin, int s_count, long long *acc) {
    return WavefrontImpl<T, Cam>::run(ctx, cam, o, rows_local, s_begin, s_count, acc);
}

template <typename T, int ACCEL> void launch_pb(const TraceArgs<T> &A, int grid, size_t smem, cudaStream_t st) {
    if constexpr (sizeof(T) == 4 || ACCEL == RT_ACCEL_LINEAR || ACCEL == RT_ACCEL_GRID) trace_kernel_pb<T, ACCEL><<<grid, TRACE_BLOCK, smem, st>>>(A);
}
template <typename T, int ACCEL> int shape_pb(rt_ctx *ctx, size_t smem, int *grid) {
    if constexpr (sizeof(T) == 4 || ACCEL == RT_ACCEL_LINEAR || ACCEL == RT_ACCEL_GRID) return launch_shape(ctx, trace_kernel_pb<T, ACCEL>, smem, grid);
    else return RT_EPRECISION;
}

// RT_ACCEL_GRID: camera rays through the tile lists, scattered rays through the uniform grid.  Float scenes of any size; double
// scenes (the walk runs in float on the rounded ray, the exact tests in double) up to 8 192 slots -- above that the tile lists
// come from a walk of the LBVH, which is a float structure.
template <typename T, typename Cam>
int trace_grid(rt_ctx *ctx, const Cam &cam, const rt_opts &o, int rows_local, int s_begin, int s_count, long long *acc) {
    const bool from_bvh = ctx->blob.n > GRID_BINS_FROM_BVH;
    if (from_bvh && sizeof(T) != 4) return RT_EPRECISION;
    int rc = build_grid(ctx);
    if (rc) return rc;
    if (!ctx->grid_usable) return RT_EINVAL;                    // not a field of similar spheres: use RT_ACCEL_LBVH
    int grid = 0;
    rc = shape_pb<T, RT_ACCEL_GRID>(ctx, 0, &grid);
    if (rc) return rc;
    TraceArgs<T> A;
    rc = fill_args<T>(ctx, cam, o, rows_local, s_begin, s_count, acc, A);
    if (rc) return rc;
    if (A.plan.total_jobs == 0) return RT_OK;
    const unsigned long long lanes = (unsigned long long)grid * TRACE_BLOCK;
    if (A.plan.total_jobs < lanes) grid = (int)((A.plan.total_jobs + TRACE_BLOCK - 1) / TRACE_BLOCK);
    ctx->stats.grid = grid;
    // large scenes: the tile lists come from a walk of the LBVH (one thread per tile cannot loop over 10^5 slots)
    if (from_bvh) { rc = build_lbvh(ctx); if (rc) return rc; }
    rc = build_bins<T>(ctx, A, cam.width, cam.height, from_bvh);
    if (rc) return rc;
    RT_CUDA(cudaMemcpyToSymbolAsync(g_grid, &ctx->grid, sizeof(GridView), 0, cudaMemcpyHostToDevice, ctx->stream));
    launch_pb<T, RT_ACCEL_GRID>(A, grid, 0, ctx->stream);
    RT_CUDA(cudaGetLastError());
    ctx->stats.launches += 1;
    return RT_OK;
}

// Launches the path tracer for samples [s_begin, s_begin + s_count) of `rows_local` rows into the accumulators `acc`
// (which the caller has zeroed).
template <typename T, typename Cam>
int trace(rt_ctx *ctx, const Cam &cam, const rt_opts &o, int rows_local, int s_begin, int s_count, long long *acc) {
    const int accel = resolve_accel(ctx, o);
    ctx->stats.accel_used = accel;
    if (o.kernel == RT_KERNEL_WAVEFRONT) return trace_wavefront<T>(ctx, cam, o, rows_local, s_begin, s_count, acc);
    if (accel == RT_ACCEL_GRID) return trace_grid<T>(ctx, cam, o, rows_local, s_begin, s_count, acc);
    const bool lbvh = (accel == RT_ACCEL_LBVH);
    if (lbvh && sizeof(T) != 4) return RT_EPRECISION;
    const size_t smem = lbvh ? 0 : trace_smem(ctx->blob);
    if (smem > 227 * 1024 || (!lbvh && ctx->blob.n > 65535)) return RT_EINVAL;   // too large for the shared-memory scan: use RT_ACCEL_LBVH
    int grid = 0;
    int rc = lbvh ? build_lbvh(ctx) : RT_OK;
    if (rc) return rc;
    const bool compact = lbvh && ctx->bvh.compact;
    // camera rays through per-tile candidate lists (rt_primary_bins.cuh): same image either way
    const bool pbins = o.primary_bins != RT_PBINS_OFF && !getenv("RT_NO_PBINS");
    if (pbins) rc = compact ? shape_pb<T, ACCEL_LBVH_COMPACT>(ctx, smem, &grid)
                            : (lbvh ? shape_pb<T, RT_ACCEL_LBVH>(ctx, smem, &grid) : shape_pb<T, RT_ACCEL_LINEAR>(ctx, smem, &grid));
    else rc = compact ? launch_shape(ctx, trace_kernel<T, ACCEL_LBVH_COMPACT>, smem, &grid)
                      : (lbvh ? launch_shape(ctx, trace_kernel<T, RT_ACCEL_LBVH>, smem, &grid)
                              : launch_shape(ctx, trace_kernel<T, RT_ACCEL_LINEAR>, smem, &grid));
    if (rc) return rc;
    TraceArgs<T> A;
    rc = fill_args<T>(ctx, cam, o, rows_local, s_begin, s_count, acc, A);
    if (rc) return rc;
    A.bvh = ctx->bvh;
    // node visits per loop turn: about one root-to-leaf descent plus slack (measured: 16 best for 487 spheres,
    // 24 for 99 860); finished lanes are shaded between rounds
    A.bvh_steps = 8;
    for (int m = ctx->bvh.m; m > 0; m >>= 1) A.bvh_steps += 1;
    if (A.bvh_steps > 32) A.bvh_steps = 32;
    A.bvh_min_active = 6;              // measured: +2.4 % on scene 1, +1.4 % on the 99 860-slot scene against 0
    if (const char *e = getenv("RT_BVH_STEPS")) A.bvh_steps = atoi(e) > 0 ? atoi(e) : A.bvh_steps;   // tuning knobs
    if (const char *e = getenv("RT_BVH_MIN_ACTIVE")) A.bvh_min_active = atoi(e);
    if (A.plan.total_jobs == 0) return RT_OK;
    const unsigned long long lanes = (unsigned long long)grid * TRACE_BLOCK;
    if (A.plan.total_jobs < lanes) grid = (int)((A.plan.total_jobs + TRACE_BLOCK - 1) / TRACE_BLOCK);
    ctx->stats.grid = grid;
    if (pbins) {
        rc = build_bins<T>(ctx, A, cam.width, cam.height, lbvh);
        if (rc) return rc;
        if (compact) launch_pb<T, ACCEL_LBVH_COMPACT>(A, grid, smem, ctx->stream);
        else if (lbvh) launch_pb<T, RT_ACCEL_LBVH>(A, grid, smem, ctx->stream);
        else launch_pb<T, RT_ACCEL_LINEAR>(A, grid, smem, ctx->stream);
    } else if (compact) trace_kernel<T, ACCEL_LBVH_COMPACT><<<grid, TRACE_BLOCK, smem, ctx->stream>>>(A);
    else if (lbvh) trace_kernel<T, RT_ACCEL_LBVH><<<grid, TRACE_BLOCK, smem, ctx->stream>>>(A);
    else trace_kernel<T, RT_ACCEL_LINEAR><<<grid, TRACE_BLOCK, smem, ctx->stream>>>(A);
    RT_CUDA(cudaGetLastError());
    ctx->stats.launches += 1;
    return RT_OK;
}

template <typename T>
int finalize(rt_ctx *ctx, const AccSources &src, unsigned long long pix, T scale, T *out_dev, RowPlacement rp = RowPlacement{0, 1, 0, 1}) {
    if (pix == 0) return RT_OK;
    if (rp.world <= 1 && (reinterpret_cast<uintptr_t>(out_dev) & (2 * sizeof(T) - 1)) == 0 && !getenv("RT_FINALIZE_BY_PIXEL")) {
        // no row placement: the frame is a flat array of 3 * pix values, each a function of its own accumulator
        const unsigned long long vals = 3ull * pix, pairs = vals >> 1;
        const unsigned grid = (unsigned)((pairs + FINALIZE_FLAT_UNROLL * 256ull - 1) / (FINALIZE_FLAT_UNROLL * 256ull));
        finalize_flat_kernel<T><<<grid ? grid : 1u, 256, 0, ctx->stream>>>(src, vals, scale, out_dev);
        RT_CUDA(cudaGetLastError());
        ctx->stats.launches += 1;
        return RT_OK;
    }
    const unsigned long long quads = (pix + 3) / 4;
    const unsigned grid = (unsigned)((quads + 255) / 256);
    finalize_kernel<T><<<grid, 256, 0, ctx->stream>>>(src, pix, scale, out_dev, rp);
    RT_CUDA(cudaGetLastError());
    ctx->stats.launches += 1;
    return RT_OK;
}

int check_opts(const rt_opts &o) {
    if (o.world < 1 || o.rank < 0 || o.rank >= o.world) return RT_EINVAL;
    if (o.split == RT_SPLIT_ROWS && o.tile_rows < 1) return RT_EINVAL;
    if (o.split < RT_SPLIT_NONE || o.split > RT_SPLIT_SPP) return RT_EINVAL;
    if (o.accel != RT_ACCEL_LINEAR && o.accel != RT_ACCEL_LBVH && o.accel != RT_ACCEL_AUTO && o.accel != RT_ACCEL_GRID) return RT_EINVAL;
    if (o.accel == RT_ACCEL_GRID && (o.kernel != RT_KERNEL_MEGA || o.primary_bins == RT_PBINS_OFF)) return RT_EINVAL;
    if (o.kernel != RT_KERNEL_MEGA && o.kernel != RT_KERNEL_WAVEFRONT) return RT_EINVAL;
    if (o.primary_bins < RT_PBINS_AUTO || o.primary_bins > RT_PBINS_ON) return RT_EINVAL;
    if (o.kernel == RT_KERNEL_WAVEFRONT && (o.accel == RT_ACCEL_LBVH || o.accel == RT_ACCEL_GRID)) return RT_EINVAL;
    return RT_OK;
}

int read_counters(rt_ctx *ctx, float ms_total, float ms_trace) {
    unsigned long long h[QUEUE_WORDS];
    RT_CUDA(cudaMemcpyAsync(h, ctx->queue, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.segments = h[1];
    ctx->stats.paths = h[2];
    ctx->stats.node_visits = h[3];
    ctx->stats.sphere_tests = h[4];
    ctx->stats.binned_segments = h[5];
    ctx->stats.filter_tests = h[6];
    ctx->stats.render_ms = ms_total;
    ctx->stats.trace_ms = ms_trace;
    ctx->stats.bvh_build_ms = ctx->bvh_build_ms;
    ctx->stats.grid_build_ms = ctx->grid_build_ms;
    return RT_OK;
}

template <typename T, typename Cam>
int render_impl(rt_ctx *ctx, const Cam *cam, const rt_opts *opts_in, T *out_rgb, float *render_ms) {
    if (!ctx || !cam || !out_rgb) return RT_EINVAL;
    if (!ctx->scene_dev) return RT_ENOSCENE;
    if (ctx->scene_prec != (int)sizeof(T)) return RT_EPRECISION;
    if (cam->width <= 0 || cam->height <= 0 || cam->spp <= 0) return RT_EINVAL;
    rt_opts o;
    if (opts_in) o = *opts_in; else rt_opts_default(&o);
    int rc = check_opts(o);
    if (rc) return rc;
    if (o.split == RT_SPLIT_SPP) return RT_EINVAL;
    RT_CUDA(cudaSetDevice(ctx->device));

    const int rows_local = (o.split == RT_SPLIT_ROWS)
                               ? rt_partition_rows(cam->height, o.tile_rows, o.rank, o.world, nullptr, 0)
                               : cam->height;
    const unsigned long long pix = (unsigned long long)rows_local * cam->width;
    ctx->stats = rt_stats{};
    const size_t out_bytes = (size_t)pix * 3 * sizeof(T);
    const bool out_on_device = is_device_ptr(out_rgb);
    const bool place = o.split == RT_SPLIT_ROWS && o.place_rows != 0;
    if (place && !out_on_device) return RT_EINVAL;                   // direct row placement needs a device frame
    RowPlacement rp{cam->width, 1, 0, 1};
    if (place) rp = RowPlacement{cam->width, o.tile_rows, o.rank, o.world};
    T *frame = out_rgb;
    if (!out_on_device) {
        rc = ensure(&ctx->frame, &ctx->frame_bytes, out_bytes ? out_bytes : 16);
        if (rc) return rc;
        frame = static_cast<T *>(ctx->frame);
    }
    if (cam->max_depth <= 0 && place) return RT_EINVAL;
    const size_t acc_bytes = (size_t)pix * 3 * sizeof(long long);
    if (cam->max_depth > 0) {
        rc = ensure(&ctx->acc, &ctx->acc_bytes, acc_bytes + 16);
        if (rc) return rc;
    }
    RT_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (cam->max_depth <= 0) {
        // GF camera.h:84,127: no bounce budget -> every path is black
        RT_CUDA(cudaMemsetAsync(frame, 0, out_bytes, ctx->stream));
        RT_CUDA(cudaMemsetAsync(ctx->queue, 0, QUEUE_WORDS * sizeof(unsigned long long), ctx->stream));
        RT_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    } else {
        RT_CUDA(cudaMemsetAsync(ctx->acc, 0, acc_bytes, ctx->stream));
        rc = trace<T>(ctx, *cam, o, rows_local, 0, cam->spp, static_cast<long long *>(ctx->acc));
        if (rc) return rc;
        RT_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
        AccSources src{};
        src.p[0] = static_cast<const long long *>(ctx->acc);
        src.n = 1;
        rc = finalize<T>(ctx, src, pix, static_cast<T>(cam->scale), frame, rp);
        if (rc) return rc;
    }
    RT_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (!out_on_device && out_bytes)
        RT_CUDA(cudaMemcpyAsync(out_rgb, frame, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(cudaEventSynchronize(ctx->ev[2]));
    float ms_total = 0.f, ms_trace = 0.f;
    RT_CUDA(cudaEventElapsedTime(&ms_total, ctx->ev[0], ctx->ev[2]));
    RT_CUDA(cudaEventElapsedTime(&ms_trace, ctx->ev[0], ctx->ev[1]));
    rc = read_counters(ctx, ms_total, ms_trace);
    if (rc) return rc;
    if (render_ms) *render_ms = ms_total;
    return RT_OK;
}

template <typename T, typename Cam>
int primary_impl(rt_ctx *ctx, const Cam *cam, int accel, int32_t *ids, T *t) {
    if (!ctx || !cam || !ids || !t) return RT_EINVAL;
    if (cam->width <= 0 || cam->height <= 0) return RT_EINVAL;
    if (accel != RT_ACCEL_LINEAR && accel != RT_ACCEL_LBVH && accel != RT_ACCEL_AUTO && accel != RT_ACCEL_GRID) return RT_EINVAL;
    if (!ctx->scene_dev) return RT_ENOSCENE;
    if (ctx->scene_prec != (int)sizeof(T)) return RT_EPRECISION;
    accel = resolve_accel(ctx, accel);
    const bool grid = accel == RT_ACCEL_GRID;
    if (accel == RT_ACCEL_LBVH && sizeof(T) != 4) return RT_EPRECISION;
    RT_CUDA(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)cam->width * cam->height;
    const bool ids_dev = is_device_ptr(ids), t_dev = is_device_ptr(t);
    int32_t *d_ids = ids;
    T *d_t = t;
    void *tmp = nullptr;
    if (!ids_dev || !t_dev) {
        RT_CUDA(cudaMalloc(&tmp, npix * (sizeof(int32_t) + sizeof(T)) + 16));
        if (!t_dev) d_t = static_cast<T *>(tmp);
        if (!ids_dev) d_ids = reinterpret_cast<int32_t *>(static_cast<char *>(tmp) + npix * sizeof(T));
    }
    const bool lbvh = accel == RT_ACCEL_LBVH;
    const size_t smem = (lbvh || grid) ? 0 : trace_smem(ctx->blob);
    int rc = RT_OK;
    if (!lbvh && !grid && (smem > 227 * 1024 || ctx->blob.n > 65535)) rc = RT_EINVAL;
    if (rc == RT_OK && lbvh) rc = build_lbvh(ctx);
    if (rc == RT_OK && grid) {
        rc = build_grid(ctx);
        if (rc == RT_OK && !ctx->grid_usable) rc = RT_EINVAL;
        if (rc == RT_OK && cudaMemcpyToSymbolAsync(g_grid, &ctx->grid, sizeof(GridView), 0, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
            rc = (int)cudaGetLastError();
    }
    if (rc != RT_OK) { if (tmp) cudaFree(tmp); return rc; }
    cudaError_t e = (lbvh || grid) ? cudaSuccess
                         : cudaFuncSetAttribute(primary_kernel<T, RT_ACCEL_LINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        int grid_dim = (int)((npix + TRACE_BLOCK - 1) / TRACE_BLOCK);
        if (grid_dim > ctx->sm_count * 8) grid_dim = ctx->sm_count * 8;
        if (grid)
            primary_kernel<T, RT_ACCEL_GRID><<<grid_dim, TRACE_BLOCK, 0, ctx->stream>>>(to_dev<T>(*cam), ctx->blob, ctx->bvh, cam->width,
                                                                                        cam->height, d_ids, d_t);
        else if (lbvh)
            primary_kernel<T, RT_ACCEL_LBVH><<<grid_dim, TRACE_BLOCK, 0, ctx->stream>>>(to_dev<T>(*cam), ctx->blob, ctx->bvh, cam->width,
                                                                                        cam->height, d_ids, d_t);
        else
            primary_kernel<T, RT_ACCEL_LINEAR><<<grid_dim, TRACE_BLOCK, smem, ctx->stream>>>(to_dev<T>(*cam), ctx->blob, ctx->bvh,
                                                                                         cam->width, cam->height, d_ids, d_t);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && !ids_dev)
        e = cudaMemcpyAsync(ids, d_ids, npix * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && !t_dev) e = cudaMemcpyAsync(t, d_t, npix * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = (int)e;
    if (tmp) cudaFree(tmp);
    return rc;
}

}  // namespace

extern "C" {

int rt_create(int device, rt_ctx **out) {
    if (!out) return RT_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return RT_ENODEVICE; }
    if (device < 0 || device >= count) return RT_EINVAL;
    RT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RT_CUDA(cudaGetDeviceProperties(&prop, device));
    // The library carries sm_100a SASS only (no PTX; "a" targets are not forward compatible): any other part, other 10.x
    // parts included, is refused here rather than at the first launch.
    if (prop.major != 10 || prop.minor != 0) return RT_ENODEVICE;
    {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, finalize_kernel<float>) != cudaSuccess) { cudaGetLastError(); return RT_ENODEVICE; }
    }
    rt_ctx *ctx = new (std::nothrow) rt_ctx;
    if (!ctx) return RT_ENOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    ctx->own_stream = (e == cudaSuccess);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&ctx->queue), QUEUE_WORDS * sizeof(unsigned long long));
    if (e != cudaSuccess) { rt_destroy(ctx); return (int)e; }
    *out = ctx;
    return RT_OK;
}

int rt_destroy(rt_ctx *ctx) {
    if (!ctx) return RT_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->scene_dev) cudaFree(ctx->scene_dev);
    if (ctx->acc) cudaFree(ctx->acc);
    if (ctx->frame) cudaFree(ctx->frame);
    if (ctx->queue) cudaFree(ctx->queue);
    for (void *m : ctx->bvh_mem) if (m) cudaFree(m);
    if (ctx->wf_mem) cudaFree(ctx->wf_mem);
    if (ctx->bins) cudaFree(ctx->bins);
    for (void *m : ctx->grid_mem) if (m) cudaFree(m);
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RT_OK;
}

int rt_set_stream(rt_ctx *ctx, void *cuda_stream) {
    if (!ctx) return RT_EINVAL;
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    return RT_OK;
}

int rt_upload_scene(rt_ctx *ctx, const rt_slot *slots, int n) { return upload<float>(ctx, slots, n); }
int rt_upload_scene64(rt_ctx *ctx, const rt_slot64 *slots, int n) { return upload<double>(ctx, slots, n); }

int rt_render(rt_ctx *ctx, const rt_camera *cam, const rt_opts *opts, float *out_rgb, float *render_ms) {
    return render_impl<float>(ctx, cam, opts, out_rgb, render_ms);
}
int rt_render64(rt_ctx *ctx, const rt_camera64 *cam, const rt_opts *opts, double *out_rgb, float *render_ms) {
    return render_impl<double>(ctx, cam, opts, out_rgb, render_ms);
}

int rt_render_partials(rt_ctx *ctx, const rt_camera *cam, const rt_opts *opts_in, int64_t *acc_dev, float *render_ms) {
    if (!ctx || !cam || !acc_dev) return RT_EINVAL;
    if (!ctx->scene_dev) return RT_ENOSCENE;
    if (ctx->scene_prec != 4) return RT_EPRECISION;
    if (cam->width <= 0 || cam->height <= 0 || cam->spp <= 0 || cam->max_depth <= 0) return RT_EINVAL;
    rt_opts o;
    if (opts_in) o = *opts_in; else rt_opts_default(&o);
    int rc = check_opts(o);
    if (rc) return rc;
    if (!is_device_ptr(acc_dev)) return RT_EINVAL;
    RT_CUDA(cudaSetDevice(ctx->device));
    int32_t s0 = 0, s1 = cam->spp;
    if (o.split == RT_SPLIT_SPP) { rc = rt_partition_samples(cam->spp, o.rank, o.world, &s0, &s1); if (rc) return rc; }
    else if (o.split == RT_SPLIT_ROWS) return RT_EINVAL;
    ctx->stats = rt_stats{};
    const size_t acc_bytes = (size_t)cam->width * cam->height * 3 * sizeof(long long);
    RT_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    RT_CUDA(cudaMemsetAsync(acc_dev, 0, acc_bytes, ctx->stream));
    rc = trace<float>(ctx, *cam, o, cam->height, s0, s1 - s0, reinterpret_cast<long long *>(acc_dev));
    if (rc) return rc;
    RT_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    RT_CUDA(cudaEventSynchronize(ctx->ev[1]));
    float ms = 0.f;
    RT_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    rc = read_counters(ctx, ms, ms);
    if (rc) return rc;
    if (render_ms) *render_ms = ms;
    return RT_OK;
}

int rt_finalize_sum(rt_ctx *ctx, const rt_camera *cam, const int64_t *const *acc_dev, int n_acc, float *out_rgb, float *finalize_ms) {
    if (!ctx || !cam || !acc_dev || !out_rgb || n_acc < 1 || n_acc > ACC_SOURCES_MAX) return RT_EINVAL;
    if (cam->width <= 0 || cam->height <= 0 || cam->spp <= 0) return RT_EINVAL;
    AccSources src{};
    for (int g = 0; g < n_acc; ++g) {
        if (!acc_dev[g] || !is_device_ptr(acc_dev[g])) return RT_EINVAL;
        src.p[g] = reinterpret_cast<const long long *>(acc_dev[g]);
    }
    src.n = n_acc;
    RT_CUDA(cudaSetDevice(ctx->device));
    const unsigned long long pix = (unsigned long long)cam->width * cam->height;
    const size_t out_bytes = (size_t)pix * 3 * sizeof(float);
    const bool out_on_device = is_device_ptr(out_rgb);
    float *frame = out_rgb;
    if (!out_on_device) {
        int rc = ensure(&ctx->frame, &ctx->frame_bytes, out_bytes);
        if (rc) return rc;
        frame = static_cast<float *>(ctx->frame);
    }
    RT_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    int rc = finalize<float>(ctx, src, pix, cam->scale, frame);
    if (rc) return rc;
    RT_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    if (!out_on_device) RT_CUDA(cudaMemcpyAsync(out_rgb, frame, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    RT_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    if (finalize_ms) *finalize_ms = ms;
    return RT_OK;
}

int rt_finalize(rt_ctx *ctx, const rt_camera *cam, const int64_t *acc_dev, float *out_rgb, float *finalize_ms) {
    return rt_finalize_sum(ctx, cam, &acc_dev, 1, out_rgb, finalize_ms);
}

int rt_primary_hits(rt_ctx *ctx, const rt_camera *cam, int32_t *ids, float *t) {
    return primary_impl<float>(ctx, cam, RT_ACCEL_LINEAR, ids, t);
}
int rt_primary_hits64(rt_ctx *ctx, const rt_camera64 *cam, int32_t *ids, double *t) {
    return primary_impl<double>(ctx, cam, RT_ACCEL_LINEAR, ids, t);
}
int rt_primary_hits_accel(rt_ctx *ctx, const rt_camera *cam, int accel, int32_t *ids, float *t) {
    return primary_impl<float>(ctx, cam, accel, ids, t);
}
int rt_primary_hits_accel64(rt_ctx *ctx, const rt_camera64 *cam, int accel, int32_t *ids, double *t) {
    return primary_impl<double>(ctx, cam, accel, ids, t);
}

int rt_frame_alloc(rt_ctx *ctx, size_t bytes, void **dev_ptr) {
    if (!ctx || !dev_ptr || bytes == 0) return RT_EINVAL;
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaMalloc(dev_ptr, bytes));
    return RT_OK;
}
int rt_frame_free(rt_ctx *ctx, void *dev_ptr) {
    if (!ctx) return RT_EINVAL;
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaFree(dev_ptr));
    return RT_OK;
}
int rt_frame_read(rt_ctx *ctx, const void *dev_ptr, void *host_ptr, size_t bytes) {
    if (!ctx || !dev_ptr || !host_ptr) return RT_EINVAL;
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaMemcpy(host_ptr, dev_ptr, bytes, cudaMemcpyDeviceToHost));
    return RT_OK;
}
int rt_frame_write(rt_ctx *ctx, void *dev_ptr, const void *host_ptr, size_t bytes) {
    if (!ctx || !dev_ptr || !host_ptr) return RT_EINVAL;
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaMemcpy(dev_ptr, host_ptr, bytes, cudaMemcpyHostToDevice));
    return RT_OK;
}
int rt_copy_peer(rt_ctx *ctx, void *dst_dev, const void *src_peer, int src_device, size_t bytes) {
    if (!ctx || !dst_dev || !src_peer) return RT_EINVAL;
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaMemcpyPeerAsync(dst_dev, ctx->device, src_peer, src_device, bytes, ctx->stream));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}
int rt_enable_peer_access(rt_ctx *ctx, int peer_device) {
    if (!ctx) return RT_EINVAL;
    if (peer_device == ctx->device) return RT_OK;
    RT_CUDA(cudaSetDevice(ctx->device));
    int can = 0;
    RT_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) return RT_ENODEVICE;
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return RT_OK; }
    return e == cudaSuccess ? RT_OK : (int)e;
}

int rt_filter_audit(rt_ctx *ctx, const rt_camera *cam, uint64_t seed, uint64_t n_rays, uint64_t out[5]) {
    if (!ctx || !cam || !out) return RT_EINVAL;
    if (!ctx->scene_dev) return RT_ENOSCENE;
    if (ctx->scene_prec != 4) return RT_EPRECISION;
    for (int q = 0; q < 5; ++q) out[q] = 0;
    if (!ctx->blob.filter_ok || n_rays == 0) return RT_OK;             // exact scan in use: nothing to audit
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaMemsetAsync(ctx->queue, 0, QUEUE_WORDS * sizeof(unsigned long long), ctx->stream));
    const unsigned long long blocks = (n_rays + 255) / 256;
    const int grid = (int)std::min<unsigned long long>(blocks, (unsigned long long)ctx->sm_count * 8);
    filter_audit_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->blob, to_dev<float>(*cam), cam->width, cam->height, philox_keys(seed), n_rays,
                                                       ctx->queue);
    RT_CUDA(cudaGetLastError());
    unsigned long long h[5];
    RT_CUDA(cudaMemcpyAsync(h, ctx->queue, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < 5; ++q) out[q] = h[q];
    return RT_OK;
}

#ifdef RT_CHECKS
namespace rt { __global__ void check_selftest_kernel() { RT_CHECK(threadIdx.x != 0, 999); } }
#endif
int rt_debug_checks(rt_ctx *ctx, int32_t *enabled, uint32_t *first_code, uint32_t *failures, int selftest) {
    if (!ctx || !enabled || !first_code || !failures) return RT_EINVAL;
    *enabled = 0; *first_code = 0; *failures = 0;
#ifdef RT_CHECKS
    RT_CUDA(cudaSetDevice(ctx->device));
    if (selftest) { rt::check_selftest_kernel<<<1, 32, 0, ctx->stream>>>(); RT_CUDA(cudaGetLastError()); }
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned int h[2] = {0u, 0u};
    const unsigned int zero[2] = {0u, 0u};
    RT_CUDA(cudaMemcpyFromSymbol(h, rt::g_rt_check, sizeof h));
    RT_CUDA(cudaMemcpyToSymbol(rt::g_rt_check, zero, sizeof zero));
    *enabled = 1; *failures = h[0]; *first_code = h[1];
#else
    (void)selftest;
#endif
    return RT_OK;
}

#ifndef RT_KERNEL_ID
#define RT_KERNEL_ID "unknown"
#endif
const char *rt_kernel_build_id(void) { return RT_KERNEL_ID; }

int rt_get_stats(rt_ctx *ctx, rt_stats *stats) {
    if (!ctx || !stats) return RT_EINVAL;
    *stats = ctx->stats;
    return RT_OK;
}

}  // extern "C"
