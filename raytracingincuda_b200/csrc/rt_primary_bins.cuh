// rt_primary_bins.cuh -- screen-space candidate bins for PRIMARY rays, and the persistent path tracer that uses them.
// Included by rt_kernels.cu after trace_kernel (it shares TraceArgs, camera_ray, scatter, sky, global_row).
//
// Why.  In the reference every ray segment pays the full hit_world scan (GF hittable.h:80-98).  A camera ray, unlike a
// scattered one, is known before the frame starts up to its two random draws (pixel jitter, lens sample; GF camera.h:
// 145-155): all camera rays of a 16 x 16 pixel tile lie inside a thin bundle around the tile's central ray.  A tiny kernel
// (bin_kernel, one thread per tile, microseconds) lists, per tile, every slot that ANY ray of that bundle can touch; the
// path tracer resolves a camera ray against its tile's list with the reference's exact arithmetic (resolve_slot) and keeps
// the shared-memory scan for the scattered segments.  Scene 1 has 2.89 segments per path, so this removes about a third of
// all scans.  The hit a camera ray gets is the one the full scan returns, bit for bit, because
//   * the closest hit is order independent (rt_device.cuh: each slot contributes its first root > tmin, the minimum wins,
//     ties go to the lowest slot), so it can be computed over any SUPERSET of the slots with a non-negative discriminant;
//   * the list is such a superset (proof below).
// A tile whose list would not fit (PB_CAP slots) is marked PB_OVERFLOW and its camera rays take the shared-memory scan.
// LBVH scenes get the same lists from bin_kernel_bvh, which walks the tree with the bundle instead of looping over the
// slots (a node's box is replaced by its bounding ball, inflated like a sphere below it), and trace_kernel_pb<float, LBVH>
// keeps the resumable traversal for the scattered segments.
//
// The bundle.  A camera ray joins a lens point L = centre + q0*disk_u + q1*disk_v, q0^2 + q1^2 < 1, to a point
// Q = pixel00 + px*du + py*dv of the focus plane with |px - i| <= 0.5, |py - j| <= 0.5.  With L0 = centre and Q0 the target
// of the tile's middle, |L - L0| <= rho (largest singular value of [disk_u disk_v]) and |Q - Q0| <= hT = hx|du| + hy|dv|
// (hx, hy: half extents of the tile in pixels).  A point of the ray is X(s) = (1-s) L + s Q, s > 0, hence
//   |X(s) - X0(s)| <= |1-s| rho + s hT <= rho + s (hT + rho),       X0(s) = (1-s) L0 + s Q0  (the axis).
// The reference's float discriminant of a sphere (c, r) can only be >= 0 when the ray passes within
//   r_eff = sqrt(r^2 + 64 * 2^-24 * D^2),  D >= |c - origin|
// of c (its rounding error is below 18 * 2^-24 * |d|^2 |c - o|^2; 64 leaves a factor 3.5).  So a
// slot can only matter if some axis point satisfies |X0(s) - c| <= R0 + kappa * s|a|, R0 = r_eff + rho, kappa = (hT + rho)/|a|,
// a = Q0 - L0.  Minimising the left side minus the right side over ALL real s (a superset of s > 0) gives the closed form
//   perp * sqrt(1 - kappa^2) - kappa * along <= R0,        along = (c - L0).a/|a|,  perp = distance of c from the axis,
// A second necessary condition removes the 2 rho this linear bound carries around the focus plane (pb_touches).  Both are
// evaluated in double with absolute and relative margins that dwarf the float rounding of camera_ray
// (|px|, |py| round monotonically, so they stay inside the tile; L and Q move by a few 1e-7 relative).
#pragma once

namespace rt {

#ifndef RT_PB_SHIFT
#define RT_PB_SHIFT 4
#endif
constexpr int PB_SHIFT = RT_PB_SHIFT;       // tiles of 16 x 16 pixels
#ifndef RT_PB_STRIDE
#define RT_PB_STRIDE 64       // 63 slots per tile: measured against 31 / 127 (profiles/logs/r02l_tile_capacity_rounds.log): 14 404 slots
#endif                        // 39.9 -> 38.0 ms (binned camera rays 0.935 -> 0.999), scene 1 and the 99 860-slot scene unchanged; 127 slower
constexpr int PB_STRIDE = RT_PB_STRIDE;     // uint32 per tile record: [0] = count or PB_OVERFLOW, [1..] = slots (a multiple of 4)
constexpr int PB_CAP = PB_STRIDE - 1;
constexpr unsigned PB_OVERFLOW = 0xffffffffu;
constexpr double PB_NOISE = 64.0 * 5.9604644775390625e-08;      // 64 * 2^-24, see above

// The bundle of a tile's camera rays: axis L0 + s*a, opening kappa, start radius rho (+ margins), all in double.
struct PbBundle {
    double L0[3], ah[3];        // lens centre, unit axis
    double kappa, cosk, rho, margin;
    double la, hT;              // |a|, half extent of the tile on the focus plane
    bool ok;                    // false: degenerate camera or a very wide tile -> PB_OVERFLOW
};

template <typename T>
__device__ __forceinline__ PbBundle pb_bundle(const DevCamera<T> &cam, int tile, int tiles_x, int width, int height) {
    PbBundle B;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int i0 = tx << PB_SHIFT, j0 = ty << PB_SHIFT;
    const int i1 = min(i0 + (1 << PB_SHIFT), width) - 1, j1 = min(j0 + (1 << PB_SHIFT), height) - 1;
    const double mx = 0.5 * (i0 + i1), my = 0.5 * (j0 + j1);
    const double hx = 0.5 * (i1 - i0) + 0.5 + 1e-3, hy = 0.5 * (j1 - j0) + 0.5 + 1e-3;
    B.L0[0] = (double)cam.center.x; B.L0[1] = (double)cam.center.y; B.L0[2] = (double)cam.center.z;
    const double du[3] = {(double)cam.du.x, (double)cam.du.y, (double)cam.du.z}, dv[3] = {(double)cam.dv.x, (double)cam.dv.y, (double)cam.dv.z};
    const double p0[3] = {(double)cam.pixel00.x, (double)cam.pixel00.y, (double)cam.pixel00.z};
    double a[3], la2 = 0.0, lq2 = 0.0, ll2 = 0.0, ndu = 0.0, ndv = 0.0;
    for (int q = 0; q < 3; ++q) {
        const double Q0 = p0[q] + mx * du[q] + my * dv[q];
        a[q] = Q0 - B.L0[q];
        la2 += a[q] * a[q]; lq2 += Q0 * Q0; ll2 += B.L0[q] * B.L0[q];
        ndu += du[q] * du[q]; ndv += dv[q] * dv[q];
    }
    const double la = sqrt(la2);
    B.rho = 0.0;
    if (!(cam.defocus_angle <= T(0))) {
        // largest singular value of the 3 x 2 matrix [disk_u disk_v]
        double uu = 0.0, vv = 0.0, uv = 0.0;
        const double U[3] = {(double)cam.disk_u.x, (double)cam.disk_u.y, (double)cam.disk_u.z};
        const double V[3] = {(double)cam.disk_v.x, (double)cam.disk_v.y, (double)cam.disk_v.z};
        for (int q = 0; q < 3; ++q) { uu += U[q] * U[q]; vv += V[q] * V[q]; uv += U[q] * V[q]; }
        B.rho = sqrt(0.5 * (uu + vv + sqrt((uu - vv) * (uu - vv) + 4.0 * uv * uv))) * (1.0 + 1e-9);
    }
    B.margin = 1e-5 * (1.0 + sqrt(ll2) + sqrt(lq2));                              // float rounding of L and Q is ~1e-7 relative
    const double hT = hx * sqrt(ndu) + hy * sqrt(ndv);
    B.kappa = (hT + B.rho + B.margin) / la * (1.0 + 1e-6);
    B.ok = (la > 0.0) && (B.kappa < 0.5) && (B.rho < 1e300) && (B.margin < 1e300);   // false for NaN/inf cameras too
    B.cosk = B.ok ? sqrt(1.0 - B.kappa * B.kappa) : 0.0;
    B.la = la;
    B.hT = hT;
    const double inv_la = B.ok ? 1.0 / la : 0.0;
    for (int q = 0; q < 3; ++q) B.ah[q] = a[q] * inv_la;
    return B;
}

// Can a ray of the bundle come within `reach(D)` of the point c?  `radius` and `rmin` describe what sits at c: a sphere
// (radius = rmin = r) or the bounding ball of a BVH box (radius = half diagonal, rmin = smallest sphere radius inside);
// the float-noise inflation sqrt(rmin^2 + PB_NOISE * D^2) - rmin is evaluated at a distance bound D that covers the whole ball.
__device__ __forceinline__ bool pb_touches(const PbBundle &B, double cx, double cy, double cz, double radius, double rmin) {
    const double b[3] = {cx - B.L0[0], cy - B.L0[1], cz - B.L0[2]};
    const double bb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
    const double along = b[0] * B.ah[0] + b[1] * B.ah[1] + b[2] * B.ah[2];
    const double perp = sqrt(fmax(bb - along * along, 0.0));
    const double D = sqrt(bb) + radius + B.rho + B.margin;
    const double reach = radius + (sqrt(rmin * rmin + PB_NOISE * D * D) - rmin);
    const double R0 = reach + B.rho + B.margin;
    const double lhs = perp * B.cosk, rhs = (R0 + B.kappa * along) * (1.0 + 1e-9) + B.margin;
    if (lhs > rhs) return false;
    // Second necessary condition, tight where the first is loose (around the focus plane the lens term of the bundle radius
    // vanishes, the linear bound above carries 2 rho there).  The first condition confines the axis points that can matter to
    // arc lengths u in [u_lo, u_hi]; the exact bundle radius w(u) = |1 - u/|a|| rho + (u/|a|) hT is convex, so it is at most
    // max(w(u_lo), w(u_hi)) there, and the ball must reach the axis within that: perp <= reach + w_max.
    const double u_lo = fmax((along - R0) / (1.0 + B.kappa), 0.0), u_hi = fmax((along + R0) / (1.0 - B.kappa), 0.0);
    const double rl = (B.rho + B.margin) * (1.0 + 1e-6), hl = (B.hT + B.margin) * (1.0 + 1e-6);
    const double s_lo = u_lo / B.la, s_hi = u_hi / B.la;
    const double w_max = fmax(fabs(1.0 - s_lo) * rl + s_lo * hl, fabs(1.0 - s_hi) * rl + s_hi * hl);
    return !(perp > (reach + w_max) * (1.0 + 1e-9) + B.margin);                   // NaN geometry counts as a candidate
}

// one thread per tile, every slot of the scene (linear-scan scenes: a few hundred to a few thousand slots)
template <typename T>
__global__ void __launch_bounds__(128) bin_kernel(const __grid_constant__ DevCamera<T> cam, const typename Num<T>::vec4 *__restrict__ geom,
                                                  int n, int width, int height, int tiles_x, int tiles_y,
                                                  unsigned int *__restrict__ bins) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= tiles_x * tiles_y) return;
    const PbBundle B = pb_bundle(cam, tile, tiles_x, width, height);
    unsigned int *rec = bins + (size_t)tile * PB_STRIDE;
    if (!B.ok) { rec[0] = PB_OVERFLOW; return; }
    int cnt = 0;
    for (int s = 0; s < n; ++s) {
        const typename Num<T>::vec4 g = geom[s];
        const double r = fabs((double)g.w);
        if (pb_touches(B, (double)g.x, (double)g.y, (double)g.z, r, r)) {
            if (cnt < PB_CAP) rec[1 + cnt] = (unsigned int)s;
            ++cnt;
        }
    }
    rec[0] = cnt <= PB_CAP ? (unsigned int)cnt : PB_OVERFLOW;
}

// one thread per tile, the bundle walks the LBVH (node records of rt_lbvh.cuh); spheres outside the tree are tested directly
__global__ void __launch_bounds__(128) bin_kernel_bvh(const __grid_constant__ DevCamera<float> cam, const __grid_constant__ BvhView bv,
                                                      int width, int height, int tiles_x, int tiles_y, unsigned int *__restrict__ bins) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= tiles_x * tiles_y) return;
    const PbBundle B = pb_bundle(cam, tile, tiles_x, width, height);
    unsigned int *rec = bins + (size_t)tile * PB_STRIDE;
    if (!B.ok) { rec[0] = PB_OVERFLOW; return; }
    int cnt = 0;
    auto sphere = [&](const float4 g, int slot) {
        const double r = fabs((double)g.w);
        if (pb_touches(B, (double)g.x, (double)g.y, (double)g.z, r, r)) {
            if (cnt < PB_CAP) rec[1 + cnt] = (unsigned int)slot;
            ++cnt;
        }
    };
    for (int b = 0; b < bv.nbig; ++b) sphere(bv.big_geom[b], bv.big_slot[b]);
    if (bv.m == 1) sphere(bv.geom[0], bv.slot[0]);
    if (bv.m > 1) {
        int stack[BVH_STACK];
        int sp = 0, node = 0;
        while (cnt <= PB_CAP) {
            const float4 q0 = bv.nodes[4 * (size_t)node], q1 = bv.nodes[4 * (size_t)node + 1];
            const float4 q2 = bv.nodes[4 * (size_t)node + 2], q3 = bv.nodes[4 * (size_t)node + 3];
            const int child[2] = {__float_as_int(q3.x), __float_as_int(q3.y)};
            const double lo[2][3] = {{q0.x, q0.y, q0.z}, {q1.z, q1.w, q2.x}}, hi[2][3] = {{q0.w, q1.x, q1.y}, {q2.y, q2.z, q2.w}};
            const double rmin[2] = {fabs((double)q3.z), fabs((double)q3.w)};
            int next = -1;
            for (int c = 0; c < 2; ++c) {
                if (child[c] < 0) { sphere(bv.geom[~child[c]], bv.slot[~child[c]]); continue; }
                const double cx = 0.5 * (lo[c][0] + hi[c][0]), cy = 0.5 * (lo[c][1] + hi[c][1]), cz = 0.5 * (lo[c][2] + hi[c][2]);
                const double ex = 0.5 * (hi[c][0] - lo[c][0]), ey = 0.5 * (hi[c][1] - lo[c][1]), ez = 0.5 * (hi[c][2] - lo[c][2]);
                // the box corners were rounded to float when the tree was built: a relative 1e-6 covers that
                const double rb = sqrt(ex * ex + ey * ey + ez * ez) * (1.0 + 1e-6) + 1e-6 * (fabs(cx) + fabs(cy) + fabs(cz));
                if (!pb_touches(B, cx, cy, cz, rb, rmin[c])) continue;
                if (next < 0) next = child[c];
                else if (sp < BVH_STACK) stack[sp++] = child[c];
                else cnt = PB_CAP + 1;                                            // cannot happen for a 30-bit Morton tree; stay safe
            }
            if (next >= 0) { node = next; continue; }
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
    rec[0] = cnt <= PB_CAP ? (unsigned int)cnt : PB_OVERFLOW;
}

// ------------------------------------------------------------------------------------------------------------------------
// Persistent path tracer with binned camera rays; the scattered segments go through the shared-memory scan
// (ACCEL = RT_ACCEL_LINEAR) or the resumable LBVH traversal (RT_ACCEL_LBVH / ACCEL_LBVH_COMPACT, float).  Same Philox
// counters per (pixel, sample) and the same integer accumulation, therefore the same image, bit for bit, as trace_kernel<T, ACCEL>.
//
// A lane is in one of four phases: FRESH (starts the next sample of its job), HIT (a closest hit waits to be shaded),
// RAY (a live ray waits for the scan / for its traversal to start), FLY (LBVH: traversal in flight).  One loop turn =
//   A  up to pb_rounds times: job fetch; FRESH lanes generate their camera ray and resolve it against the tile list --
//      a miss adds the sky and leaves the lane FRESH for the next round, a hit makes it HIT;
//   B  HIT lanes scatter (one Philox block, dimension depth+1) and become RAY, or end the path and become FRESH;
//   C  one shared-memory scan for the RAY lanes (LBVH: RAY lanes start their traversal, then a bounded number of node
//      visits for every lane in flight); misses add the sky (FRESH), hits become HIT.
// so (almost) every lane that enters the scan carries a scattered ray.
#ifndef RT_TRACE_MIN_BLOCKS_LBVH
#define RT_TRACE_MIN_BLOCKS_LBVH 4     // the traversal is latency bound: 4 CTAs of 64 registers beat 3 of 80 (99 860 slots: -7 %)
#endif
#ifndef RT_TRACE_MIN_BLOCKS_GRID
#define RT_TRACE_MIN_BLOCKS_GRID 4
#endif
#ifndef RT_COOP_UNIT
#define RT_COOP_UNIT 1                 // float kernels: warp-cooperative random_unit_vector (coop_unit_vector, rt_kernels.cu)
#endif
template <typename T, int ACCEL>
__global__ void __launch_bounds__(TRACE_BLOCK, sizeof(T) == 4 ? (ACCEL == RT_ACCEL_LINEAR ? RT_TRACE_MIN_BLOCKS : (ACCEL == RT_ACCEL_GRID ? RT_TRACE_MIN_BLOCKS_GRID : RT_TRACE_MIN_BLOCKS_LBVH)) : 2)
trace_kernel_pb(const __grid_constant__ TraceArgs<T> A) {
    using N = Num<T>;
    constexpr bool LB = (ACCEL == RT_ACCEL_LBVH || ACCEL == ACCEL_LBVH_COMPACT);
    constexpr bool RAYD = (ACCEL == ACCEL_LBVH_COMPACT);
    constexpr bool GR = (ACCEL == RT_ACCEL_GRID);                  // experimental uniform grid (rt_grid.cuh)
    static_assert(!LB || sizeof(T) == 4, "the LBVH is a float structure");
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    SceneView<T> sc;
    unsigned short *cand = nullptr;
    ScanGeom geo = scan_geom(0u, 0);
    if constexpr (!LB && !GR) {
        stage_scene(smem, A.scene.base, A.scene.bytes, &bar);
        sc = view_of<T>(smem, A.scene);
        cand = reinterpret_cast<unsigned short *>(smem + A.scene.bytes) + threadIdx.x;
        geo = scan_geom(smem_u32(smem), A.scene);
    } else {
        sc = view_of<T>(A.scene.base, A.scene);
    }
    // Work counters live in shared memory, one column per thread (they are touched once per segment / path, and five
    // registers matter in kernels that sit at their register limit): [0] segments [1] paths [2] nodes / cells
    // [3] exact sphere tests [4] binned camera rays [5] filter tests.
    __shared__ unsigned int s_cnt[6][TRACE_BLOCK];
#pragma unroll
    for (int q = 0; q < 6; ++q) s_cnt[q][threadIdx.x] = 0u;
    BvhTrav tv;
    BvhStack bvh_stack;
    tv.node = -1;

    const int lane = threadIdx.x & 31;
    enum { NEED_JOB = 0, ACTIVE = 1, DEAD = 2 };
    enum { FRESH = 0, HIT = 1, RAY = 2, FLY = 3 };
    int state = NEED_JOB, phase = FRESH;
    PathState<T> ps;
    ps.o = A.cam.center;
    ps.d.x = T(0); ps.d.y = T(1); ps.d.z = T(0);
    ps.att.x = ps.att.y = ps.att.z = T(1);
    ps.puy = T(0);
    Hit<T> hit;
    hit.t = N::inf();
    hit.id = -1;
    int pi = 0, pj = 0, sample = 0, sample_end = 0, depth = 0;
    uint32_t pixel = 0, local = 0, tile = 0;
#ifdef RT_DIAG_LONG_TRAVERSAL
    int diag_visits = 0;
#endif

    auto end_black = [&]() {
        ++s_cnt[1][threadIdx.x];
        phase = FRESH;
        if (++sample == sample_end) state = NEED_JOB;
    };
    auto end_path = [&](T cr, T cg, T cb) {
        RT_CHECK(local < A.plan.pix_local && state == ACTIVE, 401);
        accumulate<T>(A.acc, local, cr, cg, cb);                     // integer atomics: order independent (rt_device.cuh)
        end_black();
    };
    auto end_in_sky = [&]() {
        T sr, sg, sb;
        sky<T>(ps.puy, sr, sg, sb);
        end_path(N::mul(ps.att.x, sr), N::mul(ps.att.y, sg), N::mul(ps.att.z, sb));
    };
    // a segment's closest hit is known: count it, then sky or pending hit
    auto land = [&](const Hit<T> &h) {
        ++s_cnt[0][threadIdx.x];
        if (h.id < 0) end_in_sky();
        else { hit = h; phase = HIT; }
    };

    for (;;) {
        // ---- A: camera rays through the tile lists ----
#pragma unroll 1
        for (int round = 0; round < A.pb_rounds; ++round) {
            unsigned want = __ballot_sync(FULL, state == NEED_JOB);
            // cohorts: lanes that claim together get adjacent pixels (coherent rays); with pb_cohort > 1 the lanes that ran out of
            // work wait until that many of them can claim together -- unless nobody else in the warp has work left
            if (A.pb_cohort > 1 && __popc(want) < A.pb_cohort && __any_sync(FULL, state == ACTIVE)) want = 0u;
            if (want) {
                const unsigned long long claimed = claim_job(A, lane, want);
                if (state == NEED_JOB) {
                    if (claimed < A.plan.total_jobs) {
                        const JobInfo J = decode_job(A, claimed);
                        pi = J.pi; pj = J.pj; pixel = J.pixel; local = J.local;
                        sample = J.sample; sample_end = J.sample_end;
                        tile = (uint32_t)(pj >> PB_SHIFT) * (uint32_t)A.tiles_x + (uint32_t)(pi >> PB_SHIFT);
                        state = ACTIVE;
                        phase = FRESH;
                    } else {
                        state = DEAD;
                    }
                }
            }
            const bool prim = (state == ACTIVE && phase == FRESH);
            const int n_prim = __popc(__ballot_sync(FULL, prim));
            if (n_prim == 0 || (round > 0 && n_prim < A.pb_min)) break;
            if (prim) {
                Philox ph;
                ph.open(A.keys, pixel, (uint32_t)sample, 0u);
                ph.block(0);                 // outside camera_ray's candidate loop: measured 5 % faster than one Philox site in the loop
                camera_ray<T, true>(A, pi, pj, ph, ps);
                depth = 0;
                RT_CHECK(tile < (uint32_t)A.tiles, 402);
                const uint4 *rec = reinterpret_cast<const uint4 *>(A.bins) + (size_t)tile * (PB_STRIDE / 4);
                uint4 q = __ldg(rec);
                const unsigned cnt = q.x;
                RT_CHECK(cnt == PB_OVERFLOW || cnt <= (unsigned)PB_CAP, 403);
                const T a = dot3(ps.d, ps.d);
                const bool sane = a > T(1e-30) && a < T(1e30);        // false for NaN too: such rays take the scan's exact loop
                if (cnt == PB_OVERFLOW || !sane) {
                    phase = RAY;                                      // this camera ray goes through the scan / the tree
                } else {
                    Hit<T> h;
                    h.t = N::inf();
                    h.id = -1;
#pragma unroll 1
                    for (unsigned e = 1; e <= cnt; ++e) {
                        // the record is consumed one word at a time: rotate the 128-bit window, refill every 4 entries
                        if ((e & 3u) == 0u) q = __ldg(rec + (e >> 2));
                        else { q.x = q.y; q.y = q.z; q.z = q.w; }
                        RT_CHECK(q.x < (unsigned)A.scene.n, 404);
                        if constexpr (LB || GR) bvh_test_sphere<T>(ldg_geom(sc.geom + q.x), (int)q.x, ps.o, ps.d, a, h);
                        else resolve_slot<T>(geo.addr, (int)q.x, ps.o, ps.d, a, h);
                    }
                    s_cnt[3][threadIdx.x] += cnt;
                    ++s_cnt[4][threadIdx.x];
                    land(h);
                }
            }
        }
        if (__all_sync(FULL, state == DEAD)) break;

        // ---- B: shade the pending hits (GF camera.h:92-117) ----
        if constexpr (sizeof(T) == 4 && RT_COOP_UNIT) {
            // the whole warp looks for the unit vectors of the lanes whose first candidate was rejected (coop_unit_vector)
            const bool shade = (state == ACTIVE && phase == HIT);
            int type = RT_DIELECTRIC;
            T schlick_u = T(0);
            Vec3<T> cand;
            cand.x = cand.y = cand.z = T(1);
            bool need = false;
            if (shade) {
                Philox ph;
                ph.open(A.keys, pixel, (uint32_t)sample, (uint32_t)(depth + 1));
                ph.block(0);
                RT_CHECK(hit.id >= 0 && hit.id < sc.n, 302);
                type = sc.type[hit.id];
                if (type == RT_DIELECTRIC) schlick_u = N::uniform(ph.w[0], 0);
                else need = !ball_candidate(ph, cand);
            }
            coop_unit_vector(A.keys, need, pixel, (uint32_t)sample, (uint32_t)(depth + 1), cand);
            if (shade) {
                const bool alive = scatter_with<T>(sc, hit, type, schlick_u, cand, ps);
                if (!alive || ++depth >= A.max_depth) end_black();                // GF camera.h:117 / :84,127 -> black
                else phase = RAY;
            }
        } else if (state == ACTIVE && phase == HIT) {
            Philox ph;
            ph.open(A.keys, pixel, (uint32_t)sample, (uint32_t)(depth + 1));
            ph.block(0);
            const bool alive = scatter<T, true>(sc, hit, ph, ps);
            if (!alive || ++depth >= A.max_depth) end_black();                    // GF camera.h:117 / :84,127 -> black
            else phase = RAY;
        }

        // ---- C: closest hit of the scattered rays ----
        if constexpr (LB) {
            unsigned int n_nodes = 0, n_tests = 0;
            if (state == ACTIVE && phase == RAY) {
                bvh_start<RAYD>(A.bvh, ps.o, ps.d, tv, n_tests);
                phase = FLY;
#ifdef RT_DIAG_LONG_TRAVERSAL
                diag_visits = 0;
#endif
            }
#pragma unroll 1
            for (int step = 0; step < A.bvh_steps; ++step) {
                const int flying_lanes = __popc(__ballot_sync(FULL, tv.node >= 0));
                if (flying_lanes == 0 || (flying_lanes < A.bvh_min_active && step > 0)) break;
                if (tv.node >= 0) bvh_step<RAYD>(A.bvh, ps.o, ps.d, tv, bvh_stack, n_nodes, n_tests);
#ifdef RT_DIAG_LONG_TRAVERSAL
                // diagnostic build only: which rays make very long traversals?
                if (tv.node >= 0 && ++diag_visits % RT_DIAG_LONG_TRAVERSAL == 0)
                    printf("long traversal: %d visits pixel (%d,%d) sample %d depth %d o %.9g %.9g %.9g d %.9g %.9g %.9g best t %g id %d sp %d\n", diag_visits,
                           pi, pj, sample, depth, (double)ps.o.x, (double)ps.o.y, (double)ps.o.z, (double)ps.d.x, (double)ps.d.y, (double)ps.d.z,
                           (double)tv.hit.t, tv.hit.id, tv.sp);
#endif
            }
            s_cnt[2][threadIdx.x] += n_nodes;
            s_cnt[3][threadIdx.x] += n_tests;
            if (state == ACTIVE && phase == FLY && tv.node < 0) land(tv.hit);
        } else if constexpr (GR) {
            // a grid walk is short (1.3 cells on average, tools/analyse_accel.py): it runs to the end right here
            if (state == ACTIVE && phase == RAY) {
                unsigned int n_nodes = 0, n_tests = 0;
                const Hit<T> h = grid_closest_hit<T>(g_grid, sc.geom, ps.o, ps.d, n_nodes, n_tests);
                s_cnt[2][threadIdx.x] += n_nodes;
                s_cnt[3][threadIdx.x] += n_tests;
                land(h);
            }
        } else {
            // all 32 lanes take part in the shared-memory scan; the ones without a live ray scan a stale one
            const bool scan = (state == ACTIVE && phase == RAY);
            if (__any_sync(FULL, scan)) {
                ScanCount cnt{0u, 0u};
                const Hit<T> h = closest_hit<T>(geo, A.scene.n, ps.o, ps.d, cand, TRACE_BLOCK, cnt);
                s_cnt[3][threadIdx.x] += cnt.exact;
                s_cnt[5][threadIdx.x] += cnt.filt;
                if (scan) land(h);
            }
        }
    }

    flush_counters(A.queue, lane, s_cnt[0][threadIdx.x], s_cnt[1][threadIdx.x], s_cnt[2][threadIdx.x], s_cnt[3][threadIdx.x],
                   s_cnt[4][threadIdx.x], s_cnt[5][threadIdx.x]);
}

}  // namespace rt
