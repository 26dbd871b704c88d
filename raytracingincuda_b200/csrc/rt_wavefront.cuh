// rt_wavefront.cuh -- the wavefront variant of the render loop (north star (c): "a wavefront
// variant that sorts by material to cut divergence").  Included by rt_kernels.cu after the
// shared building blocks (TraceArgs, PathState, camera_ray, scatter, sky, closest_hit).
//
// Same state machine as trace_kernel -- a pool slot owns one (pixel, sample range) job at a time, every
// (pixel, sample) path is the same pure function of the seed and the radiance goes into the same integer
// accumulators, so the image is bit-identical to the megakernel's -- but the state lives in a global SoA
// pool and each loop turn is two kernels:
//
//   wf_shade      runs over the slots SORTED BY WHAT THEY NEED: [camera ray | lambertian |
//                 metal | dielectric]; each class is padded to a warp multiple, so every warp
//                 executes one branch of the material switch (GF camera.h:92-108).
//   wf_intersect  closest hit for every live slot (the same shared-memory scan), ends missed
//                 paths with the sky term, and appends each slot to the class list of the next
//                 turn with one warp-aggregated atomic per class (counting sort by material).
#pragma once

namespace rt {

enum { WF_NEED_JOB = 0, WF_FRESH = 1, WF_HIT = 2, WF_DEAD = 3 };
enum { WF_CLASSES = 4 };                      // 0 camera ray (fresh / needs job), 1 + material type

struct WfPool {
    // SoA, `n` slots each
    float *ox, *oy, *oz, *dx, *dy, *dz, *ax, *ay, *az, *puy, *hit_t;
    int *hit_id, *sample, *sample_end, *depth, *state;
    uint32_t *pixel, *local;
    int n;
    int *list_in, *list_out;                  // [WF_CLASSES][n] slot indices
    unsigned int *count_in, *count_out;       // [WF_CLASSES]
    unsigned int *alive;                      // live slots seen by the last wf_intersect
};

__device__ __forceinline__ void wf_load_path(const WfPool &P, int s, PathState<float> &ps) {
    ps.o.x = P.ox[s]; ps.o.y = P.oy[s]; ps.o.z = P.oz[s];
    ps.d.x = P.dx[s]; ps.d.y = P.dy[s]; ps.d.z = P.dz[s];
    ps.att.x = P.ax[s]; ps.att.y = P.ay[s]; ps.att.z = P.az[s];
    ps.puy = P.puy[s];
}
__device__ __forceinline__ void wf_store_path(const WfPool &P, int s, const PathState<float> &ps) {
    P.ox[s] = ps.o.x; P.oy[s] = ps.o.y; P.oz[s] = ps.o.z;
    P.dx[s] = ps.d.x; P.dy[s] = ps.d.y; P.dz[s] = ps.d.z;
    P.ax[s] = ps.att.x; P.ay[s] = ps.att.y; P.az[s] = ps.att.z;
    P.puy[s] = ps.puy;
}

// a path of slot `s` ended with radiance c: accumulate, advance the job (same as trace_kernel's end_path)
__device__ __forceinline__ int wf_end_path(const TraceArgs<float> &A, const WfPool &P, int s, float cr, float cg, float cb,
                                           unsigned int &n_path) {
    RT_CHECK(s >= 0 && s < P.n && P.local[s] < A.plan.pix_local, 701);
    accumulate<float>(A.acc, P.local[s], cr, cg, cb);
    ++n_path;
    const int smp = P.sample[s] + 1;
    P.sample[s] = smp;
    return smp == P.sample_end[s] ? WF_NEED_JOB : WF_FRESH;
}

// Position of thread `i` in the class-sorted order: classes are padded to warp multiples.
__device__ __forceinline__ int wf_slot_of(const WfPool &P, unsigned int i, int &cls) {
    unsigned int base = 0;
#pragma unroll
    for (int c = 0; c < WF_CLASSES; ++c) {
        const unsigned int cnt = P.count_in[c], padded = (cnt + 31u) & ~31u;
        if (i < base + padded) {
            cls = c;
            return (i - base) < cnt ? P.list_in[(size_t)c * P.n + (i - base)] : -1;
        }
        base += padded;
    }
    cls = -1;
    return -1;
}

__global__ void __launch_bounds__(256) wf_shade(const __grid_constant__ TraceArgs<float> A, const __grid_constant__ WfPool P) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    // materials come from the staged scene blob like in the megakernel
    stage_scene(smem, A.scene.base, A.scene.bytes, &bar);
    const SceneView<float> sc = view_of<float>(smem, A.scene);
    const int lane = threadIdx.x & 31;
    int cls;
    const int s = wf_slot_of(P, blockIdx.x * blockDim.x + threadIdx.x, cls);
    if (cls < 0) return;                                   // whole warp is past the last class
    unsigned int n_path = 0;
    int state = s >= 0 ? P.state[s] : WF_DEAD;
    PathState<float> ps;
    uint32_t pixel = 0;
    int sample = 0, depth = 0;
    if (s >= 0) { pixel = P.pixel[s]; sample = P.sample[s]; depth = P.depth[s]; }

    if (cls > 0 && s >= 0) {
        // ---- material classes: shade the pending hit (one branch per warp) ----
        wf_load_path(P, s, ps);
        Hit<float> hit;
        hit.t = P.hit_t[s];
        hit.id = P.hit_id[s];
        Philox ph;
        ph.open(A.keys, pixel, (uint32_t)sample, (uint32_t)(depth + 1));
        ph.block(0);
        const bool alive = scatter(sc, hit, ph, ps);
        if (!alive || ++depth >= A.max_depth) {
            state = wf_end_path(A, P, s, 0.0f, 0.0f, 0.0f, n_path);
            sample += 1;
        } else {
            state = WF_HIT;                                // has a new ray; wf_intersect decides what comes next
        }
    }
    // ---- class 0 (and the rare path that just ended above): job fetch + camera ray ----
    const unsigned want = __ballot_sync(0xffffffffu, s >= 0 && state == WF_NEED_JOB);
    if (want) {
        const int leader = __ffs(want) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(A.queue, (unsigned long long)__popc(want));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (s >= 0 && state == WF_NEED_JOB) {
            const unsigned long long job = base + (unsigned long long)__popc(want & ((1u << lane) - 1u));
            if (job < A.plan.total_jobs) {
                const JobInfo J = decode_job(A, job);
                pixel = J.pixel;
                sample = J.sample;
                P.local[s] = J.local;
                P.pixel[s] = pixel;
                P.sample[s] = sample;
                P.sample_end[s] = J.sample_end;
                state = WF_FRESH;
            } else {
                state = WF_DEAD;
            }
        }
    }
    if (s >= 0 && state == WF_FRESH) {
        Philox ph;
        ph.open(A.keys, pixel, (uint32_t)sample, 0u);
        ph.block(0);
        camera_ray(A, (int)(pixel % (uint32_t)A.width), (int)(pixel / (uint32_t)A.width), ph, ps);
        depth = 0;
        state = WF_HIT;                                    // "has a ray"
    }
    if (s >= 0) {
        if (state == WF_HIT) { wf_store_path(P, s, ps); P.depth[s] = depth; }
        P.state[s] = state;
    }
    unsigned long long pth = n_path;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) pth += __shfl_xor_sync(0xffffffffu, pth, off);
    if (lane == 0 && pth) atomicAdd(A.queue + 2, pth);
}

__global__ void __launch_bounds__(256) wf_intersect(const __grid_constant__ TraceArgs<float> A, const __grid_constant__ WfPool P) {
    using N = Num<float>;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    stage_scene(smem, A.scene.base, A.scene.bytes, &bar);
    const SceneView<float> sc = view_of<float>(smem, A.scene);
    unsigned short *cand = reinterpret_cast<unsigned short *>(smem + A.scene.bytes) + threadIdx.x;
    const ScanGeom geo = scan_geom(smem_u32(smem), A.scene);      // the paired filter scan of the megakernel
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = s < P.n && P.state[s] == WF_HIT;
    if (!__any_sync(0xffffffffu, live)) return;
    Vec3<float> o, d;
    o.x = o.y = o.z = 0.0f;
    d.x = 0.0f; d.y = 1.0f; d.z = 0.0f;
    if (live) { o.x = P.ox[s]; o.y = P.oy[s]; o.z = P.oz[s]; d.x = P.dx[s]; d.y = P.dy[s]; d.z = P.dz[s]; }
    ScanCount cnt{0u, 0u};
    const Hit<float> hit = closest_hit<float>(geo, A.scene.n, o, d, cand, 256, cnt);
    int cls = -1;
    unsigned int n_path = 0;
    if (live) {
        if (hit.id < 0) {
            float sr, sg, sb;
            sky<float>(P.puy[s], sr, sg, sb);
            const int st = wf_end_path(A, P, s, N::mul(P.ax[s], sr), N::mul(P.ay[s], sg), N::mul(P.az[s], sb), n_path);
            P.state[s] = st;
            cls = 0;
        } else {
            P.hit_t[s] = hit.t;
            P.hit_id[s] = hit.id;
            cls = 1 + sc.type[hit.id];
        }
    }
    // counting sort by class: one atomic per warp and class
#pragma unroll
    for (int c = 0; c < WF_CLASSES; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, cls == c);
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned int base = 0;
            if (lane == leader) base = atomicAdd(P.count_out + c, (unsigned int)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (cls == c) P.list_out[(size_t)c * P.n + base + __popc(m & ((1u << lane) - 1u))] = s;
        }
    }
    unsigned long long seg = live ? 1ull : 0ull, pth = n_path, ext = cnt.exact, flt = cnt.filt;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        seg += __shfl_xor_sync(0xffffffffu, seg, off);
        pth += __shfl_xor_sync(0xffffffffu, pth, off);
        ext += __shfl_xor_sync(0xffffffffu, ext, off);
        flt += __shfl_xor_sync(0xffffffffu, flt, off);
    }
    if (lane == 0) {
        atomicAdd(A.queue + 1, seg);
        if (pth) atomicAdd(A.queue + 2, pth);
        atomicAdd(A.queue + 4, ext);
        atomicAdd(A.queue + 6, flt);
        atomicAdd(P.alive, (unsigned int)seg);
    }
}

// every slot starts in class 0 needing a job
__global__ void wf_init(const __grid_constant__ WfPool P) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < P.n) { P.state[s] = WF_NEED_JOB; P.list_in[s] = s; }
    if (s == 0) {
        P.count_in[0] = (unsigned int)P.n;
        for (int c = 1; c < WF_CLASSES; ++c) P.count_in[c] = 0;
        for (int c = 0; c < WF_CLASSES; ++c) P.count_out[c] = 0;
        *P.alive = 0;
    }
}

}  // namespace rt
