// rt_device.cuh -- device-side building blocks of the B200 render path: explicit-rounding
// arithmetic, Philox4x32-10, the TMA bulk scene stage, and the closest-hit scan.
//
// Arithmetic contract.  The reference's results are fixed by where ptxas
// fused multiplies and adds in its render kernel (sm_100 SASS of GF hittable.h:40-66:
// FADD x3, FMUL+FFMA+FFMA for h, FMUL+FFMA+FFMA for |oc|^2, FFMA(-r,r,q), FMUL a*c,
// FFMA(h,h,-m)).  Here every operation that matters is spelled with a round-to-nearest intrinsic
// so no compiler version can move a rounding; the file is also compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt {

// ------------------------------------------------------------------------------------------
// Bounds checks of the CHECKED build (librt_b200_checked.so, -DRT_CHECKS=1).  compute-sanitizer is closed on the GPU pool
// this was developed on, so the memcheck role is played by assertions of our own at every indexed access of the kernels
// (accumulators, tile lists, candidate words, scene slots, BVH nodes and stack, grid cells, job decode); a violation is
// counted in g_rt_check[0] and its code kept in g_rt_check[1] (rt_debug_checks reads and resets them), the access is still
// made.  The production build compiles the macro away: its SASS is unaffected.
#ifdef RT_CHECKS
__device__ unsigned int g_rt_check[2];
#define RT_CHECK(cond, code)                                                              \
    do {                                                                                  \
        if (!(cond)) { if (atomicAdd(&rt::g_rt_check[0], 1u) == 0u) rt::g_rt_check[1] = (unsigned)(code); } \
    } while (0)
#else
#define RT_CHECK(cond, code) do { } while (0)
#endif

template <typename T> struct Num;

template <> struct Num<float> {
    using vec4 = float4;
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float rcp(float a) { return __frcp_rn(a); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ float min(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float tmin() { return 0.001f; }            // (float)0.001, GF camera.h:87
    static __device__ __forceinline__ float near_zero() { return 1e-6f; }        // GF vec3.h:50
    static __device__ __forceinline__ float unit_min() { return 1e-8f; }         // GF vec3.h:124
    static constexpr int words_per_uniform = 1;
    // curand_uniform: x * 2^-32 + 2^-33, in (0,1]  (curand_uniform.h:69-72)
    static __device__ __forceinline__ float uniform(uint32_t x, uint32_t) {
        return __fmaf_rn(__uint2float_rn(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    }
};

template <> struct Num<double> {
    using vec4 = double4;
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double rcp(double a) { return __drcp_rn(a); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
    static __device__ __forceinline__ double min(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double tmin() { return 0.001; }
    static __device__ __forceinline__ double near_zero() { return 1e-8; }        // GD vec3.h:50
    static __device__ __forceinline__ double unit_min() { return 1e-160; }       // GD vec3.h:125
    static constexpr int words_per_uniform = 2;
    // curand_uniform_double for 2 words: (x ^ (y << 21)) * 2^-53 + 2^-54  (curand_uniform.h:101-106)
    static __device__ __forceinline__ double uniform(uint32_t x, uint32_t y) {
        const unsigned long long z = (unsigned long long)x ^ ((unsigned long long)y << 21);
        return __dadd_rn(__dmul_rn(__ull2double_rn(z), 1.1102230246251565e-16), 1.1102230246251565e-16 / 2.0);
    }
};

template <typename T> struct Vec3 { T x, y, z; };

// dot() of the reference as compiled: fma(u.z, v.z, fma(u.x, v.x, u.y * v.y))  (GF vec3.h:93-97)
template <typename T>
__device__ __forceinline__ T dot3(const Vec3<T> &u, const Vec3<T> &v) {
    using N = Num<T>;
    return N::fma(u.z, v.z, N::fma(u.x, v.x, N::mul(u.y, v.y)));
}

// ------------------------------------------------------------------------------------------
// Order-independent accumulation.  The radiance of one path-sample is converted to 64-bit fixed
// point, round(L * 2^40), and ADDED with an integer atomic to the pixel's three accumulators.  Integer addition is
// associative, so the pixel sum -- and with it the image -- does not depend on which lane, CTA, job partition or GPU traced
// which sample, nor on the order in which they finished: the canonical image is
//     gamma( scale * float( sum_s fix40(L(p, s)) * 2^-40 ) )
// with no partial planes and no ordered reduction.  L <= 1 for the reference's materials (albedo <= 1, sky <= 1), so the
// sum stays below spp * 2^40 (exact in int64 up to 2^23 samples); multiplying by 2^40 is exact in float and double, the
// conversion rounds to nearest-even and saturates, and a NaN sample counts as 0.
constexpr int FIX_SHIFT = 40;
__device__ __forceinline__ long long fix_of(float v) {
    return (v == v) ? __float2ll_rn(__fmul_rn(v, 1099511627776.0f)) : 0ll;
}
__device__ __forceinline__ long long fix_of(double v) {
    return (v == v) ? __double2ll_rn(__dmul_rn(v, 1099511627776.0)) : 0ll;
}
// acc[3 * pixel + k] += fix40(c_k); black samples add nothing
template <typename T>
__device__ __forceinline__ void accumulate(long long *__restrict__ acc, uint32_t local_pixel, T cr, T cg, T cb) {
    unsigned long long *a = reinterpret_cast<unsigned long long *>(acc) + 3ull * local_pixel;
    const long long fr = fix_of(cr), fg = fix_of(cg), fb = fix_of(cb);
    if (fr) atomicAdd(a, (unsigned long long)fr);
    if (fg) atomicAdd(a + 1, (unsigned long long)fg);
    if (fb) atomicAdd(a + 2, (unsigned long long)fb);
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10, counter = (pixel, sample, dimension, block), key = seed.
// The ten round keys depend on the seed only: the host expands them once (philox_keys) and the
// kernels read them from the parameter constant bank, so a block is 20 IMAD.WIDE + 20 LOP3.
struct PhiloxKeys { uint32_t k[20]; };      // k[2r], k[2r+1] = key of round r

inline PhiloxKeys philox_keys(uint64_t seed) {
    PhiloxKeys K;
    uint32_t ka = (uint32_t)seed, kb = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { K.k[2 * r] = ka; K.k[2 * r + 1] = kb; ka += 0x9E3779B9u; kb += 0xBB67AE85u; }
    return K;
}

struct Philox {
    uint32_t c0, c1, c2;              // c3 (block) is supplied per call
    const uint32_t *rk;               // round keys (constant bank)
    uint32_t w[4];
    __device__ __forceinline__ void open(const PhiloxKeys &keys, uint32_t pixel, uint32_t sample, uint32_t dim) {
        c0 = pixel; c1 = sample; c2 = dim; rk = keys.k;
    }
    __device__ __forceinline__ void block(uint32_t blk) {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = blk;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            const uint32_t y0 = hi1 ^ x1 ^ rk[2 * r], y2 = hi0 ^ x3 ^ rk[2 * r + 1];
            x0 = y0; x1 = lo1; x2 = y2; x3 = lo0;
        }
        w[0] = x0; w[1] = x1; w[2] = x2; w[3] = x3;
    }
};

// ------------------------------------------------------------------------------------------
// Scene blob as uploaded by rt_upload_scene: one contiguous, 16-byte aligned device buffer
//   [vec4 geom[n8]]      centre.xyz, radius; n8 = n rounded up to 8, padding is zero records
//                        (the only array the scan reads)
//   [vec4 matl[n]]       albedo.xyz, param (fuzz for metal, refraction index for dielectric)
//   [int  type[n4]]
//   [T    rinv[n]]       1/radius (host IEEE division, the value rcp.rn gives on the device)
//   [float4 filt[2*half_pad]]  centre.xyz (rounded to float), -(|centre|^2 - radius^2) per slot, two halves
//   [int  far[n_far]]          slots kept out of filt[]
// staged into shared memory by one thread with cp.async.bulk + an mbarrier (TMA bulk copy).
template <typename T> struct SceneView {
    const typename Num<T>::vec4 *geom;
    const typename Num<T>::vec4 *matl;
    const int *type;
    const T *rinv;
    int n;
};

struct SceneBlob {
    const void *base;       // device pointer
    uint32_t bytes;         // multiple of 16
    uint32_t matl_off, type_off, rinv_off;
    int n;
    // conservative pre-filter of the float scan (see "paired filter scan" below); filter_ok == 0: exact scan only
    uint32_t filt_off, far_off;
    int filter_ok;
    int n_half;             // slots [0, n_half) live in half 0 of filt[], slots [n_half, n) in half 1
    int half_pad;           // records per half (n_half rounded up to 8; padding and far slots never pass)
    int n_far;              // slots excluded from the filter (far from the origin): always tested exactly
    float bound;            // max over filtered slots of |centre| + radius, rounded up
};

template <typename T>
__device__ __forceinline__ SceneView<T> view_of(const void *base, const SceneBlob &b) {
    SceneView<T> v;
    const char *p = static_cast<const char *>(base);
    v.geom = reinterpret_cast<const typename Num<T>::vec4 *>(p);
    v.matl = reinterpret_cast<const typename Num<T>::vec4 *>(p + b.matl_off);
    v.type = reinterpret_cast<const int *>(p + b.type_off);
    v.rinv = reinterpret_cast<const T *>(p + b.rinv_off);
    v.n = b.n;
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// global -> shared TMA bulk copy of `bytes` (multiple of 16), completion on `bar`.
// Must be called by ALL threads of the CTA (it contains the CTA barriers).
__device__ __forceinline__ void stage_scene(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                            uint64_t *bar) {
    const uint32_t bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        uint32_t off = 0;
        while (off < bytes) {
            const uint32_t piece = bytes - off < 32768u ? bytes - off : 32768u;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32(static_cast<char *>(smem_dst) + off)),
                "l"(static_cast<const char *>(gmem_src) + off), "r"(piece), "r"(bar_a)
                : "memory");
            off += piece;
        }
    }
    // every thread waits for phase 0 of the barrier (try_wait suspends the thread in hardware)
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "RT_STAGE_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra RT_STAGE_DONE;\n\t"
        "bra RT_STAGE_WAIT;\n\t"
        "RT_STAGE_DONE:\n\t"
        "}" ::"r"(bar_a)
        : "memory");
}

// ------------------------------------------------------------------------------------------
// Explicit shared-space loads of one geometry record (centre.xyz, radius).  A 32-bit shared
// address + immediate offset keeps the scan at one LDS.128 per sphere with no address math.
template <typename T> __device__ __forceinline__ typename Num<T>::vec4 lds_geom(uint32_t addr);
template <> __device__ __forceinline__ float4 lds_geom<float>(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
template <> __device__ __forceinline__ double4 lds_geom<double>(uint32_t addr) {
    double4 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(v.z), "=d"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t sign_word(float x) { return __float_as_uint(x); }
__device__ __forceinline__ uint32_t sign_word(double x) { return (uint32_t)__double2hiint(x); }

// ------------------------------------------------------------------------------------------
// hit_world (GF hittable.h:80-98) as a two-phase scan over geometry in shared memory.
//  phase 1: slots are visited in blocks of 32.  Per slot: one LDS.128, the discriminant exactly
//           as the reference computes it (12 FP32 instructions) and ONE funnel shift that
//           collects its sign bit -- no compare, no branch.  After a block the 32 sign bits are
//           inspected once; slots with disc >= 0 (0.4 % of tests) are appended to a per-thread
//           candidate list in shared memory.
//           (disc is never -0: fma(h,h,-m) of an exact cancellation rounds to +0, so "sign bit
//           clear" is exactly the reference's !(disc < 0) for every non-NaN value; a NaN
//           discriminant yields no hit in the reference either, GF hittable.h:49-55.)
//  phase 2: candidates are revisited in slot order with the reference's root logic and a
//           shrinking tmax, so (slot id, t) is what the reference's loop produces: strict
//           tmin < t < closest, lowest slot wins ties.
// `cand` points at this thread's column of a [CAND_CAP][blockDim.x] uint16 array.
// After the full blocks a tail of up to three groups of 8 covers n % 32; the geometry array is
// padded with zero records to a multiple of 8 and `tail_mask` clears the padding's bits.
constexpr int CAND_CAP = 32;       // exact scan: a list of 32 uint16 slots per thread; the paired filter scan uses the same
                                   // 64 bytes per thread as 16 candidate-mask words (PAIR_CHUNK)

template <typename T> struct Hit { T t; int id; };

// Work a lane actually executed in the scan (for the executed-FP32 roofline, bench.py): filter tests (7 FMA each) and exact
// sphere tests (the reference's 12 FP32 instructions each, plus sqrt/div on the few that pass).
struct ScanCount { unsigned int filt, exact; };

struct ScanGeom {
    uint32_t addr;        // shared-space byte address of geom[0]
    int blocks;           // n / 32 full blocks
    int tail_groups;      // ceil((n % 32) / 8) groups of 8 after the full blocks
    uint32_t tail_mask;   // valid-slot bits of the tail word (slot k of a word <-> bit 31-k)
    // paired filter scan (blob.filter_ok)
    uint32_t filt_addr, far_addr;
    int filter_ok, n_half, half_pad, n_far;
    float bound;
};

__device__ __forceinline__ ScanGeom scan_geom(uint32_t addr, int n) {
    ScanGeom g;
    g.addr = addr;
    g.blocks = n >> 5;
    const int rem = n & 31;
    g.tail_groups = (rem + 7) >> 3;
    g.tail_mask = rem ? ~(0xffffffffu >> rem) : 0u;
    g.filt_addr = g.far_addr = 0u;
    g.filter_ok = g.n_half = g.half_pad = g.n_far = 0;
    g.bound = 0.0f;
    return g;
}

// scan description of a scene blob staged at shared-space address `addr`
__device__ __forceinline__ ScanGeom scan_geom(uint32_t addr, const SceneBlob &b) {
    ScanGeom g = scan_geom(addr, b.n);
    g.filt_addr = addr + b.filt_off;
    g.far_addr = addr + b.far_off;
    g.filter_ok = b.filter_ok;
    g.n_half = b.n_half;
    g.half_pad = b.half_pad;
    g.n_far = b.n_far;
    g.bound = b.bound;
    return g;
}

template <typename T>
__device__ __forceinline__ T disc_of(const typename Num<T>::vec4 &s, const Vec3<T> &o, const Vec3<T> &d, T a, T &h) {
    using N = Num<T>;
    const T ocx = N::sub(s.x, o.x), ocy = N::sub(s.y, o.y), ocz = N::sub(s.z, o.z);
    h = N::fma(ocz, d.z, N::fma(ocx, d.x, N::mul(ocy, d.y)));
    const T q = N::fma(ocz, ocz, N::fma(ocx, ocx, N::mul(ocy, ocy)));
    const T c = N::fma(-s.w, s.w, q);
    return N::fma(h, h, -N::mul(a, c));
}

// GF hittable.h:49-56 for one slot whose discriminant is >= 0: nearest root strictly inside
// (tmin, hit.t), else the far root, else no hit.
template <typename T>
__device__ __forceinline__ void try_slot(uint32_t geom_addr, int id, const Vec3<T> &o, const Vec3<T> &d, T a,
                                         T disc, T h, Hit<T> &hit) {
    using N = Num<T>;
    (void)geom_addr; (void)o; (void)d;
    const T sq = N::sqrt(disc);
    T root = N::div(N::sub(h, sq), a);
    if (!(N::tmin() < root && root < hit.t)) {
        root = N::div(N::add(h, sq), a);
        if (!(N::tmin() < root && root < hit.t)) return;
    }
    hit.t = root;
    hit.id = id;
}

// The reference's loop verbatim (GF hittable.h:86-92): used when a ray has more candidate slots
// than the per-thread list holds.  Kept out of line: it is cold and the kernel is I-cache sensitive.
template <typename T>
__device__ __noinline__ Hit<T> rescan_in_order(uint32_t geom_addr, int n, Vec3<T> o, Vec3<T> d, T a) {
    using N = Num<T>;
    constexpr uint32_t REC = sizeof(typename N::vec4);
    Hit<T> hit;
    hit.t = N::inf();
    hit.id = -1;
    for (int id = 0; id < n; ++id) {
        const typename N::vec4 s = lds_geom<T>(geom_addr + (uint32_t)id * REC);
        T h;
        const T disc = disc_of<T>(s, o, d, a, h);
        if (!(disc < T(0))) try_slot<T>(geom_addr, id, o, d, a, disc, h, hit);
    }
    return hit;
}

// Appends the slots flagged in `m` (slot k of the word <-> bit 31-k) to the candidate list.
__device__ __forceinline__ void push_candidates(uint32_t m, int base, unsigned short *cand, int stride, int &count) {
    do {
        const int k = __clz(m);
        m &= ~(0x80000000u >> k);
        if (count < CAND_CAP) cand[count * stride] = static_cast<unsigned short>(base + k);
        ++count;
    } while (m);
}

#ifndef RT_SCAN_UNROLL
#define RT_SCAN_UNROLL 16          // slots per unrolled body; 32 / RT_SCAN_UNROLL bodies per sign word
#endif

template <typename T>
__device__ __forceinline__ Hit<T> closest_hit_exact(const ScanGeom &g, int n, const Vec3<T> &o, const Vec3<T> &d,
                                                    unsigned short *cand, int stride, ScanCount &cnt) {
    using N = Num<T>;
    constexpr uint32_t REC = sizeof(typename N::vec4);
    const T a = dot3(d, d);                                     // GF hittable.h:42
    int count = 0;
    uint32_t addr = g.addr;
    // full blocks of 32 slots: one sign word each
    for (int b = 0; b < g.blocks; ++b) {
        uint32_t signs = 0;
#pragma unroll 1
        for (int part = 0; part < 32 / RT_SCAN_UNROLL; ++part, addr += RT_SCAN_UNROLL * REC) {
#pragma unroll
            for (int k = 0; k < RT_SCAN_UNROLL; ++k) {
                const typename N::vec4 s = lds_geom<T>(addr + (uint32_t)k * REC);
                T h;
                const T disc = disc_of<T>(s, o, d, a, h);           // GF hittable.h:41-46
                signs = __funnelshift_l(sign_word(disc), signs, 1); // signs = signs << 1 | (disc < 0)
            }
        }
        const uint32_t m = ~signs;
        if (m) push_candidates(m, b * 32, cand, stride, count);    // GF hittable.h:47, 0.4 % of tests
    }
    // tail: up to three groups of 8 slots (the array is padded with zero records to a multiple of 8)
    if (g.tail_groups) {
        uint32_t signs = 0;
#pragma unroll 1
        for (int part = 0; part < g.tail_groups; ++part, addr += 8u * REC) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const typename N::vec4 s = lds_geom<T>(addr + (uint32_t)k * REC);
                T h;
                const T disc = disc_of<T>(s, o, d, a, h);
                signs = __funnelshift_l(sign_word(disc), signs, 1);
            }
        }
        const uint32_t m = (~signs << (32 - 8 * g.tail_groups)) & g.tail_mask;
        if (m) push_candidates(m, g.blocks * 32, cand, stride, count);
    }
    Hit<T> hit;
    hit.t = N::inf();
    hit.id = -1;
    cnt.exact += (unsigned)(32 * g.blocks + 8 * g.tail_groups) + (unsigned)(count <= CAND_CAP ? count : n);
    if (count <= CAND_CAP) {
#pragma unroll 1
        for (int k = 0; k < count; ++k) {
            const int id = cand[k * stride];
            const typename N::vec4 s = lds_geom<T>(g.addr + (uint32_t)id * REC);
            T h;
            const T disc = disc_of<T>(s, o, d, a, h);
            try_slot<T>(g.addr, id, o, d, a, disc, h, hit);
        }
    } else {
        // more candidate slots than the list holds (never in the reference's scenes): rescan
        // every slot in order, exactly like the reference's loop
        hit = rescan_in_order<T>(g.addr, n, o, d, a);
    }
    return hit;
}

// ------------------------------------------------------------------------------------------
// Paired filter scan.  Same answer as closest_hit_exact, about half the issue slots.
//
// (1) Conservative filter.  With d' = d/|d| the reference's discriminant has the sign of
//         F = (c.d' - o.d')^2 + 2 c.o - (|c|^2 - r^2) - |o|^2          ( = r^2 - dist(centre, ray)^2 )
//     which needs 7 fused multiply-adds per (ray, sphere) against a per-ray threshold when
//     nk = -(|c|^2 - r^2) is precomputed per sphere:  h = fma(cx,dx', fma(cy,dy', fma(cz,dz', -o.d')));
//     t = fma(cx,2ox, fma(cy,2oy, fma(cz,2oz, nk)));  v = fma(h,h,t);  candidate iff v >= |o|^2 - E.
//     E = 72 * 2^-24 * (|o| + bound)^2 covers BOTH the rounding error of v and the rounding error of
//     the reference's own float discriminant (derivation: DESIGN.md section 6), so every slot whose
//     reference discriminant is >= 0 is a candidate; candidates then go through the reference's exact
//     arithmetic (disc_of, IEEE sqrt/div), which also rejects the filter's false positives.
//     Spheres far from the origin relative to the rest (the 1000-unit ground sphere) would inflate E for
//     every slot; they are kept out of filt[] and tested exactly for every ray (blob.far).
// (2) Two rays per lane.  Lanes 2i and 2i+1 exchange their ray constants; each scans HALF of the
//     slots for BOTH rays, so one LDS.128 feeds two tests (the LSU return path, not the FMA pipe, bounds
//     a one-ray-per-load scan on sm_100) and the two tests run as packed FFMA2 with the sphere scalars
//     broadcast.  Per sphere: 1 LDS.128 + 7 FFMA2 + 2 FSETP + 2 predicated OR.
// (3) The closest hit is order independent (each sphere contributes its first root > tmin, the
//     minimum wins, ties go to the lowest slot: same rule as the LBVH leaves), so own-half, other-half
//     and far candidates can be resolved in any order.
// Must be called by all 32 lanes of the warp.
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // 1 MUFU; within 2 ulp
    return y;
}

// Are both roots of a sphere test at or below tmin?  Decided WITHOUT the IEEE square root and divisions, conservatively:
// true only when the reference's own far root, div.rn(h + sqrt.rn(disc), a), is certainly <= tmin (then the near root is
// too, and GF hittable.h:49-56 accepts nothing).  This is the commonest test of all -- every scattered ray is tested
// against the sphere it just left (c ~ 0, h < 0: roots ~ 2h/a and ~ 0) -- and it was paying ~45 instructions of
// sqrt + 2 divisions to reject.  sqrt.approx is within 2 ulp, so s_hi = approx * (1 + 4e-7) >= sqrt.rn(disc); rounding is
// monotonic, so fl(h + s_hi) >= fl(h + sqrt.rn(disc)); and x <= tmin * a * (1 - 1e-6) (two roundings of 2^-24 each) implies
// x / a < tmin, hence div.rn(x, a) <= tmin.  (disc below 1e-30 is left to the exact path: the .ftz approximation flushes
// subnormal inputs to zero.)
__device__ __forceinline__ bool roots_below_tmin(float h, float disc, float a) {
    const float s_hi = sqrt_approx(disc) * 1.0000004f;
    return disc >= 1e-30f && (h + s_hi) <= (Num<float>::tmin() * a) * 0.999999f;
}
__device__ __forceinline__ bool roots_below_tmin(double, double, double) { return false; }

template <typename T>
__device__ __forceinline__ void resolve_slot(uint32_t geom_addr, int id, const Vec3<T> &o, const Vec3<T> &d, T a, Hit<T> &hit) {
    using N = Num<T>;
    RT_CHECK(id >= 0, 101);
    const typename N::vec4 s = lds_geom<T>(geom_addr + (uint32_t)id * (uint32_t)sizeof(typename N::vec4));
    const T ocx = N::sub(s.x, o.x), ocy = N::sub(s.y, o.y), ocz = N::sub(s.z, o.z);
    const T h = N::fma(ocz, d.z, N::fma(ocx, d.x, N::mul(ocy, d.y)));
    const T q = N::fma(ocz, ocz, N::fma(ocx, ocx, N::mul(ocy, ocy)));
    const T c = N::fma(-s.w, s.w, q);
    // sphere behind the origin (h < 0, origin outside): m = a*c >= 0, so sqrt(disc) <= |h| and both roots are <= 0 < tmin
    // in the reference's own arithmetic too -- skip the square root and the divisions
    if (h < T(0) && c > T(0)) return;
    const T disc = N::fma(h, h, -N::mul(a, c));                  // GF hittable.h:41-46
    if (disc < T(0)) return;                                     // GF hittable.h:47 (also drops filter false positives)
    if (roots_below_tmin(h, disc, a)) return;                    // e.g. the sphere the ray just left
    const T sq = N::sqrt(disc);
    T v = N::div(N::sub(h, sq), a);
    if (!(N::tmin() < v)) {
        v = N::div(N::add(h, sq), a);
        if (!(N::tmin() < v)) return;
    }
    if (v < hit.t || (v == hit.t && id < hit.id)) { hit.t = v; hit.id = id; }
}

#ifndef RT_PAIR_UNROLL
#define RT_PAIR_UNROLL 16           // records per unrolled body (= per candidate-mask word) of the paired scan: 8 or 16
#endif
static_assert(RT_PAIR_UNROLL == 8 || RT_PAIR_UNROLL == 16, "a block's two candidate masks share one 32-bit word");
constexpr int PAIR_CHUNK = 16;       // blocks per chunk: one 16-bit flag word per ray, [PAIR_CHUNK][blockDim] mask words
constexpr float RT_FILTER_K = 72.0f * 5.9604644775390625e-08f;      // 72 * 2^-24

struct PairRay { float2 dx, dy, dz, ox, oy, oz, nod; float thr_own, thr_nb; };

// per-ray constants of the filter (any rounding here is inside the 72 * 2^-24 budget)
struct FilterRay { float a, dx, dy, dz, ox2, oy2, oz2, nod, thr; bool sane; };

__device__ __forceinline__ FilterRay filter_ray(const Vec3<float> &o, const Vec3<float> &d, float bound) {
    using N = Num<float>;
    FilterRay f;
    f.a = dot3(d, d);                                              // GF hittable.h:42
    const float inv = N::rcp(N::sqrt(f.a));
    f.dx = N::mul(d.x, inv); f.dy = N::mul(d.y, inv); f.dz = N::mul(d.z, inv);
    f.nod = -N::fma(o.x, f.dx, N::fma(o.y, f.dy, N::mul(o.z, f.dz)));
    const float oo = N::fma(o.x, o.x, N::fma(o.y, o.y, N::mul(o.z, o.z)));
    const float reach = N::add(N::sqrt(oo), bound);
    f.thr = N::fma(-N::mul(RT_FILTER_K, reach), reach, oo);
    f.ox2 = N::add(o.x, o.x); f.oy2 = N::add(o.y, o.y); f.oz2 = N::add(o.z, o.z);
    f.sane = f.a > 1e-30f && f.a < 1e30f && oo < 1e24f;           // false for NaN/inf too: such rays take the exact loop
    return f;
}

// the filter value of one filt[] record, scalar form (each component of the packed FFMA2 chain in
// filter_pair rounds exactly like these fmaf)
__device__ __forceinline__ float filter_value(const float4 q, const FilterRay &f) {
    using N = Num<float>;
    const float h = N::fma(q.x, f.dx, N::fma(q.y, f.dy, N::fma(q.z, f.dz, f.nod)));
    const float t = N::fma(q.x, f.ox2, N::fma(q.y, f.oy2, N::fma(q.z, f.oz2, q.w)));
    return N::fma(h, h, t);
}

// one filt[] record against both rays; bit `bit` of s_own / s_nb is set when the slot is a candidate
__device__ __forceinline__ void filter_pair(const float4 q, const PairRay &r, uint32_t bit_own, uint32_t bit_nb, uint32_t &s_own,
                                            uint32_t &s_nb) {
    const float2 cx = make_float2(q.x, q.x), cy = make_float2(q.y, q.y), cz = make_float2(q.z, q.z), nk = make_float2(q.w, q.w);
    float2 h = __ffma2_rn(cz, r.dz, r.nod);
    float2 t = __ffma2_rn(cz, r.oz, nk);
    h = __ffma2_rn(cy, r.dy, h);
    t = __ffma2_rn(cy, r.oy, t);
    h = __ffma2_rn(cx, r.dx, h);
    t = __ffma2_rn(cx, r.ox, t);
    const float2 v = __ffma2_rn(h, h, t);
    asm("{ .reg .pred p; setp.ge.f32 p, %1, %2; @p or.b32 %0, %0, %3; }" : "+r"(s_own) : "f"(v.x), "f"(r.thr_own), "r"(bit_own));
    asm("{ .reg .pred p; setp.ge.f32 p, %1, %2; @p or.b32 %0, %0, %3; }" : "+r"(s_nb) : "f"(v.y), "f"(r.thr_nb), "r"(bit_nb));
}

//
// Double scenes run the SAME float filter on float-rounded copies of the ray and of the geometry: rounding the
// inputs moves r^2 - dist^2 by less than the reference's float discriminant error the budget already holds
// (DESIGN.md section 6), and the candidates are resolved in double.
template <typename T>
__device__ __forceinline__ Hit<T> closest_hit_paired(const ScanGeom &g, int n, const Vec3<T> &o, const Vec3<T> &d,
                                                     unsigned short *cand, int stride, ScanCount &cnt) {
    using N = Num<T>;
    constexpr unsigned FULLMASK = 0xffffffffu;
    Vec3<float> of, df;
    of.x = (float)o.x; of.y = (float)o.y; of.z = (float)o.z;
    df.x = (float)d.x; df.y = (float)d.y; df.z = (float)d.z;
    const FilterRay fr = filter_ray(of, df, g.bound);
    const T a = sizeof(T) == 4 ? (T)fr.a : dot3(d, d);             // GF hittable.h:42
    const float dx = fr.dx, dy = fr.dy, dz = fr.dz, nod = fr.nod, thr = fr.thr;
    const bool sane = fr.sane;
    PairRay r;
    r.dx = make_float2(dx, __shfl_xor_sync(FULLMASK, dx, 1));
    r.dy = make_float2(dy, __shfl_xor_sync(FULLMASK, dy, 1));
    r.dz = make_float2(dz, __shfl_xor_sync(FULLMASK, dz, 1));
    const float ox2 = fr.ox2, oy2 = fr.oy2, oz2 = fr.oz2;
    r.ox = make_float2(ox2, __shfl_xor_sync(FULLMASK, ox2, 1));
    r.oy = make_float2(oy2, __shfl_xor_sync(FULLMASK, oy2, 1));
    r.oz = make_float2(oz2, __shfl_xor_sync(FULLMASK, oz2, 1));
    r.nod = make_float2(nod, __shfl_xor_sync(FULLMASK, nod, 1));
    r.thr_own = thr;
    r.thr_nb = __shfl_xor_sync(FULLMASK, thr, 1);

    // Candidate bookkeeping without a branch in the scan: after every block of 16 (or 8) records the two 16-bit
    // candidate masks go to this lane's word column in shared memory with one store, and one bit per ray says
    // whether the block had any candidate.  Blocks are handled in chunks of PAIR_CHUNK (one flag word per ray); the
    // reference's scenes are a single chunk.
    const int half = threadIdx.x & 1;
    uint32_t *words = reinterpret_cast<uint32_t *>(cand - threadIdx.x) + threadIdx.x;   // [PAIR_CHUNK][stride] uint32, my column
    const uint32_t *peer_words = words + (half ? -1 : 1);
    const int wstride = stride;                                      // cand points at column threadIdx.x of uint32 words
    Hit<T> hit;
    hit.t = N::inf();
    hit.id = -1;
    const int slot_own0 = half ? g.n_half : 0, slot_peer0 = half ? 0 : g.n_half;   // slot of record 0: my half / the peer's half
    uint32_t addr = g.filt_addr + (uint32_t)(half * g.half_pad) * 16u;
    int k0 = 0;                                                      // record index within the half
    while (k0 < g.half_pad) {
        const int chunk0 = k0;
        uint32_t flags_own = 0, flags_nb = 0;
        int blk = 0;
#pragma unroll 1
        for (; blk < PAIR_CHUNK && k0 + RT_PAIR_UNROLL <= g.half_pad; ++blk, k0 += RT_PAIR_UNROLL, addr += RT_PAIR_UNROLL * 16u) {
            uint32_t s_own = 0, s_nb = 0;
            RT_CHECK(blk < PAIR_CHUNK && k0 + RT_PAIR_UNROLL <= g.half_pad && addr + RT_PAIR_UNROLL * 16u <= g.filt_addr + 2u * (uint32_t)g.half_pad * 16u, 104);
#pragma unroll
            for (int k = 0; k < RT_PAIR_UNROLL; ++k)
                filter_pair(lds_geom<float>(addr + (uint32_t)k * 16u), r, 1u << k, 0x10000u << k, s_own, s_nb);
            words[blk * wstride] = s_own | s_nb;
            flags_own |= (s_own ? 1u : 0u) << blk;
            flags_nb |= (s_nb ? 1u : 0u) << blk;
        }
#if RT_PAIR_UNROLL > 8
        // the last block of a half may hold a single group of 8 records (half_pad is a multiple of 8)
        if (blk < PAIR_CHUNK && k0 < g.half_pad && g.half_pad - k0 < RT_PAIR_UNROLL) {
            uint32_t s_own = 0, s_nb = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) filter_pair(lds_geom<float>(addr + (uint32_t)k * 16u), r, 1u << k, 0x10000u << k, s_own, s_nb);
            words[blk * wstride] = s_own | s_nb;
            flags_own |= (s_own ? 1u : 0u) << blk;
            flags_nb |= (s_nb ? 1u : 0u) << blk;
            k0 = g.half_pad;
        }
#endif
        __syncwarp();                                                // the neighbour's words are read below
        uint32_t fo = flags_own, fp = __shfl_xor_sync(FULLMASK, flags_nb, 1);   // fp: blocks of the peer's half with candidates for MY ray
        if (sane) {
            uint32_t w = 0;
            int base = 0;
#pragma unroll 1
            while (w | fo | fp) {
                if (!w) {                                            // next block that holds a candidate for my ray
                    if (fo) {
                        const int b = __ffs(fo) - 1;
                        fo &= fo - 1u;
                        w = words[b * wstride] & 0xffffu;
                        base = slot_own0 + chunk0 + b * RT_PAIR_UNROLL;
                    } else {
                        const int b = __ffs(fp) - 1;
                        fp &= fp - 1u;
                        w = peer_words[b * wstride] >> 16;
                        base = slot_peer0 + chunk0 + b * RT_PAIR_UNROLL;
                    }
                }
                const int k = __ffs(w) - 1;
                w &= w - 1u;
                RT_CHECK(base + k < n && k < RT_PAIR_UNROLL, 102);                   // padding records (nk = -inf) never pass the filter
                resolve_slot<T>(g.addr, base + k, o, d, a, hit);
                ++cnt.exact;
            }
        }
        __syncwarp();                                                // words are rewritten by the next chunk / scan
    }
    cnt.filt += 2u * (unsigned)g.half_pad;                           // every lane: its half of the records x two rays
    if (sane) {
#pragma unroll 1
        for (int k = 0; k < g.n_far; ++k) {
            int id;
            asm volatile("ld.shared.s32 %0, [%1];" : "=r"(id) : "r"(g.far_addr + (uint32_t)k * 4u));
            RT_CHECK(id >= 0 && id < n, 103);
            resolve_slot<T>(g.addr, id, o, d, a, hit);
        }
        cnt.exact += (unsigned)g.n_far;
    } else {
        // Degenerate ray (inf / denormal scale): the reference's loop, slot by slot.  A ray with a NaN component needs no loop:
        // it poisons h or |oc|^2 of every slot, every discriminant is NaN, and GF hittable.h:47-55 accepts nothing (all its
        // comparisons are false).  Such rays are not exotic: a hit on the reference's never-written zero-radius slot has the
        // normal (p - c) * (1/0); in scene 1 about one segment in 10^4 is one, and its 488-slot loop stalled the whole warp
        // (1.4 % of the stall samples before this shortcut).
        const bool nan_ray = !(o.x == o.x && o.y == o.y && o.z == o.z && d.x == d.x && d.y == d.y && d.z == d.z);
        if (!nan_ray) { hit = rescan_in_order<T>(g.addr, n, o, d, a); cnt.exact += (unsigned)n; }
    }
    return hit;
}

// hit_world (GF hittable.h:80-98).  All 32 lanes of the warp must call it together.
template <typename T>
__device__ __forceinline__ Hit<T> closest_hit(const ScanGeom &g, int n, const Vec3<T> &o, const Vec3<T> &d,
                                              unsigned short *cand, int stride, ScanCount &cnt) {
    if (g.filter_ok) return closest_hit_paired<T>(g, n, o, d, cand, stride, cnt);
    return closest_hit_exact<T>(g, n, o, d, cand, stride, cnt);
}
template <typename T>
__device__ __forceinline__ Hit<T> closest_hit(const ScanGeom &g, int n, const Vec3<T> &o, const Vec3<T> &d,
                                              unsigned short *cand, int stride) {
    ScanCount cnt{0u, 0u};
    return closest_hit<T>(g, n, o, d, cand, stride, cnt);
}

}  // namespace rt
