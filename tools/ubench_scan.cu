// ubench_scan.cu -- issue/pipe model of the sphere-scan inner loop on sm_100a.
// Measures SM clocks per (ray, sphere) test and per warp for several formulations of the
// "does this sphere need an exact test" filter, at the occupancy of trace_kernel (4 x 256 / SM).
//   v14   the exact discriminant (3 FADD 3 FMUL 6 FFMA) + LDS.128 + SHF           (round-1 scan)
//   v10   conservative filter, scalar: 8 FFMA + LDS.128 + SHF
//   p8    conservative filter, packed: 7 FFMA2 + FADD2 per PAIR, 2 LDS.128, 2 SHF
//   p7    conservative filter, packed: 7 FFMA2 per pair, 2 LDS.128, 2 FSETP, 2 predicated OR
//   hB    half-precision projected-distance filter: 8 HFMA2 + HSET2 + LOP3 per PAIR of slots (one ray per lane), LDS.128 per pair
//   hA    the same with two rays per lane (slot scalars broadcast by operand swizzle), LDS.128 per two slots = four tests
//   ffma / ffma2 / hfma2 / hfma2v  pure FMA streams (pipe peak check; hfma2v = three varying register operands)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_scan ubench_scan.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int NS = 480;            // slots (multiple of 32)
constexpr int REPS = 256;           // scans per thread

__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(256, 4) scan_kernel(const float4 *scene, float *out, unsigned *cnt, long long *clk) {
    const long long t0 = clock64();
    __shared__ float4 s[NS];
    for (int i = threadIdx.x; i < NS; i += blockDim.x) s[i] = scene[i];
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(s);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float ox = 13.f + 1e-3f * (tid & 1023), oy = 2.f + 1e-3f * (tid & 511), oz = 3.f - 1e-3f * (tid & 127);
    float dx = -0.9f + 1e-4f * (tid & 255), dy = -0.1f - 1e-4f * (tid & 63), dz = -0.3f + 1e-4f * (tid >> 8);
    if ((MODE >= 1 && MODE <= 3) || (MODE >= 6 && MODE <= 9)) { const float il = rsqrtf(dx * dx + dy * dy + dz * dz); dx *= il; dy *= il; dz *= il; }
    unsigned found = 0;
    float accum = 0.f;
    for (int rep = 0; rep < REPS; ++rep) {
        if (MODE == 0) {
            const float a = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            for (int b = 0; b < NS / 32; ++b) {
                uint32_t signs = 0;
#pragma unroll 1
                for (int part = 0; part < 2; ++part) {
                    const uint32_t addr = base + (b * 32 + part * 16) * 16;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float4 q = lds4(addr + k * 16);
                        const float ocx = __fsub_rn(q.x, ox), ocy = __fsub_rn(q.y, oy), ocz = __fsub_rn(q.z, oz);
                        const float h = __fmaf_rn(ocz, dz, __fmaf_rn(ocx, dx, __fmul_rn(ocy, dy)));
                        const float qq = __fmaf_rn(ocz, ocz, __fmaf_rn(ocx, ocx, __fmul_rn(ocy, ocy)));
                        const float c = __fmaf_rn(-q.w, q.w, qq);
                        const float disc = __fmaf_rn(h, h, -__fmul_rn(a, c));
                        signs = __funnelshift_l(__float_as_uint(disc), signs, 1);
                    }
                }
                if (~signs) found += __popc(~signs);
            }
        } else if (MODE == 1) {
            const float nod = -__fmaf_rn(oz, dz, __fmaf_rn(ox, dx, oy * dy));
            const float ox2 = 2.f * ox, oy2 = 2.f * oy, oz2 = 2.f * oz;
            const float nthr = -(ox * ox + oy * oy + oz * oz) + 1e-3f;
            for (int b = 0; b < NS / 32; ++b) {
                uint32_t signs = 0;
#pragma unroll 1
                for (int part = 0; part < 2; ++part) {
                    const uint32_t addr = base + (b * 32 + part * 16) * 16;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float4 q = lds4(addr + k * 16);
                        const float h = __fmaf_rn(q.x, dx, __fmaf_rn(q.y, dy, __fmaf_rn(q.z, dz, nod)));
                        const float t = __fmaf_rn(q.x, ox2, __fmaf_rn(q.y, oy2, __fmaf_rn(q.z, oz2, q.w)));
                        const float v = __fmaf_rn(h, h, t);
                        const float w = __fadd_rn(v, nthr);
                        signs = __funnelshift_l(__float_as_uint(w), signs, 1);
                    }
                }
                if (~signs) found += __popc(~signs);
            }
        } else if (MODE == 2 || MODE == 3) {
            const float nod = -__fmaf_rn(oz, dz, __fmaf_rn(ox, dx, oy * dy));
            const float2 dx2 = make_float2(dx, dx), dy2 = make_float2(dy, dy), dz2 = make_float2(dz, dz);
            const float2 ox2 = make_float2(2.f * ox, 2.f * ox), oy2 = make_float2(2.f * oy, 2.f * oy), oz2 = make_float2(2.f * oz, 2.f * oz);
            const float2 nod2 = make_float2(nod, nod);
            const float thr = (ox * ox + oy * oy + oz * oz) - 1e-3f;
            const float2 nthr2 = make_float2(-thr, -thr);
            for (int b = 0; b < NS / 32; ++b) {
                uint32_t signs = 0;
#pragma unroll 1
                for (int part = 0; part < 2; ++part) {
                    const uint32_t addr = base + (b * 32 + part * 16) * 16;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 A = lds4(addr + k * 32);          // cx_a cx_b cy_a cy_b
                        const float4 B = lds4(addr + k * 32 + 16);     // cz_a cz_b nk_a nk_b
                        const float2 cx = make_float2(A.x, A.y), cy = make_float2(A.z, A.w), cz = make_float2(B.x, B.y), nk = make_float2(B.z, B.w);
                        float2 h = __ffma2_rn(cz, dz2, nod2);
                        h = __ffma2_rn(cy, dy2, h);
                        h = __ffma2_rn(cx, dx2, h);
                        float2 t = __ffma2_rn(cz, oz2, nk);
                        t = __ffma2_rn(cy, oy2, t);
                        t = __ffma2_rn(cx, ox2, t);
                        const float2 v = __ffma2_rn(h, h, t);
                        if (MODE == 2) {
                            const float2 w = __fadd2_rn(v, nthr2);
                            signs = __funnelshift_l(__float_as_uint(w.x), signs, 1);
                            signs = __funnelshift_l(__float_as_uint(w.y), signs, 1);
                        } else {
                            asm("{ .reg .pred p; setp.ge.f32 p, %1, %2; @p or.b32 %0, %0, %3; }" : "+r"(signs) : "f"(v.x), "f"(thr), "r"(1u << (2 * k)));
                            asm("{ .reg .pred p; setp.ge.f32 p, %1, %2; @p or.b32 %0, %0, %3; }" : "+r"(signs) : "f"(v.y), "f"(thr), "r"(1u << (2 * k + 1)));
                        }
                    }
                    if (MODE == 3) { if (signs) found += __popc(signs); signs = 0; }
                }
                if (MODE == 2) { if (~signs) found += __popc(~signs); }
            }
        } else if (MODE == 6 || MODE == 7 || MODE == 8 || MODE == 9) {
            // two rays per lane, one sphere per LDS.128: 7 FFMA2 per sphere with the sphere scalars broadcast
            const float ex = ox + 0.37f, ey = oy + 0.11f, ez = oz - 0.23f;                 // second ray
            const float fx = dz, fy = dy, fz = dx;
            const float nodA = -__fmaf_rn(oz, dz, __fmaf_rn(ox, dx, oy * dy)), nodB = -__fmaf_rn(ez, fz, __fmaf_rn(ex, fx, ey * fy));
            const float2 dx2 = make_float2(dx, fx), dy2 = make_float2(dy, fy), dz2 = make_float2(dz, fz);
            const float2 ox2 = make_float2(2.f * ox, 2.f * ex), oy2 = make_float2(2.f * oy, 2.f * ey), oz2 = make_float2(2.f * oz, 2.f * ez);
            const float2 nod2 = make_float2(nodA, nodB);
            const float thrA = (ox * ox + oy * oy + oz * oz) - 1e-3f, thrB = (ex * ex + ey * ey + ez * ez) - 1e-3f;
            const float2 nthr2 = make_float2(-thrA, -thrB);
            constexpr int U = (MODE == 8) ? 32 : 16;
            for (int b = 0; b < NS / 2 / U; ++b) {                 // each lane scans half of the slots for two rays
                uint32_t sA = 0, sB = 0;
                const uint32_t addr = base + (b * U) * 16;
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const float4 q = lds4(addr + k * 16);
                    const float2 cx = make_float2(q.x, q.x), cy = make_float2(q.y, q.y), cz = make_float2(q.z, q.z), nk = make_float2(q.w, q.w);
                    float2 h = __ffma2_rn(cz, dz2, nod2);
                    h = __ffma2_rn(cy, dy2, h);
                    h = __ffma2_rn(cx, dx2, h);
                    float2 t = __ffma2_rn(cz, oz2, nk);
                    t = __ffma2_rn(cy, oy2, t);
                    t = __ffma2_rn(cx, ox2, t);
                    if (MODE == 9) { h.x = fmaxf(h.x, 0.f); h.y = fmaxf(h.y, 0.f); }
                    const float2 v = __ffma2_rn(h, h, t);
                    if (MODE == 6 || MODE == 8) {
                        const float2 w = __fadd2_rn(v, nthr2);
                        sA = __funnelshift_l(__float_as_uint(w.x), sA, 1);
                        sB = __funnelshift_l(__float_as_uint(w.y), sB, 1);
                    } else {
                        asm("{ .reg .pred p; setp.ge.f32 p, %1, %2; @p or.b32 %0, %0, %3; }" : "+r"(sA) : "f"(v.x), "f"(thrA), "r"(1u << k));
                        asm("{ .reg .pred p; setp.ge.f32 p, %1, %2; @p or.b32 %0, %0, %3; }" : "+r"(sB) : "f"(v.y), "f"(thrB), "r"(1u << k));
                    }
                }
                if (MODE == 6) { sA = ~sA & 0xffffu; sB = ~sB & 0xffffu; }
                if (MODE == 8) { sA = ~sA; sB = ~sB; }
                if (sA | sB) found += __popc(sA) + __popc(sB);
            }
        } else if (MODE == 10 || MODE == 11) {
            // projected distance in half2: a = c.u - o.u, b = c.v - o.v, s = a^2 + b^2 - r^2 <= thr
            const __half2 ux = __floats2half2_rn(dy, MODE == 11 ? dz : dy), uy = __floats2half2_rn(-dx, MODE == 11 ? 0.1f : -dx), uz = __floats2half2_rn(0.f, MODE == 11 ? -dx : 0.f);
            const __half2 vx = __floats2half2_rn(dz * dx, dz * dx), vy = __floats2half2_rn(dz * dy, dz * dy), vz = __floats2half2_rn(-dx * dx - dy * dy, -dx * dx - dy * dy);
            const __half2 nou = __floats2half2_rn(-(ox * dy - oy * dx), -(ox * dy - oy * dx) + (MODE == 11 ? 0.3f : 0.f));
            const __half2 nov = __floats2half2_rn(-(ox * dz * dx + oy * dz * dy), -(ox * dz * dx + oy * dz * dy));
            const __half2 thr = __floats2half2_rn(0.02f, 0.02f);
            const uint32_t hbase = base;
            constexpr int TESTS_PER_LDS = (MODE == 10) ? 2 : 4;
            for (int b = 0; b < NS / 2 / 16; ++b) {                       // 16 tests per ray and block (MODE 11: each lane scans half of the slots)
                uint32_t acc = 0;
#pragma unroll
                for (int k = 0; k < 32 / TESTS_PER_LDS; ++k) {
                    uint4 q;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(hbase + (b * (32 / TESTS_PER_LDS) + k) * 16));
#pragma unroll
                    for (int j = 0; j < (MODE == 10 ? 1 : 2); ++j) {
                        __half2 cx, cy, cz, nr2;
                        if (MODE == 10) {
                            cx = *reinterpret_cast<__half2 *>(&q.x); cy = *reinterpret_cast<__half2 *>(&q.y);
                            cz = *reinterpret_cast<__half2 *>(&q.z); nr2 = *reinterpret_cast<__half2 *>(&q.w);
                        } else {
                            const __half2 xy = *reinterpret_cast<__half2 *>(j ? &q.z : &q.x), zr = *reinterpret_cast<__half2 *>(j ? &q.w : &q.y);
                            cx = __low2half2(xy); cy = __high2half2(xy); cz = __low2half2(zr); nr2 = __high2half2(zr);
                        }
                        __half2 a = __hfma2(cz, uz, nou); a = __hfma2(cy, uy, a); a = __hfma2(cx, ux, a);
                        __half2 bb = __hfma2(cz, vz, nov); bb = __hfma2(cy, vy, bb); bb = __hfma2(cx, vx, bb);
                        __half2 sres = __hfma2(a, a, nr2); sres = __hfma2(bb, bb, sres);
                        const unsigned m = __hle2_mask(sres, thr);
                        acc |= m & (0x00010001u << ((MODE == 10) ? k : (2 * k + j)));
                    }
                }
                if (acc) found += __popc(acc);
            }
        } else if (MODE == 12 || MODE == 13) {
            __half2 x0 = __floats2half2_rn(ox, oy), x1 = __floats2half2_rn(oz, dx), x2 = __floats2half2_rn(dy, dz), x3 = __floats2half2_rn(ox + 1.f, oy + 1.f);
            __half2 x4 = __floats2half2_rn(oy, ox), x5 = __floats2half2_rn(dx, oz), x6 = __floats2half2_rn(dz, dy), x7 = __floats2half2_rn(ox - 1.f, oy - 1.f);
            const __half2 m = __floats2half2_rn(dx, dx), c = __floats2half2_rn(dy, dz);
#pragma unroll 1
            for (int it = 0; it < NS; ++it) {
                if (MODE == 12) {
                    x0 = __hfma2(x0, m, c); x1 = __hfma2(x1, m, c); x2 = __hfma2(x2, m, c); x3 = __hfma2(x3, m, c);
                    x4 = __hfma2(x4, m, c); x5 = __hfma2(x5, m, c); x6 = __hfma2(x6, m, c); x7 = __hfma2(x7, m, c);
                } else {
                    x0 = __hfma2(x0, x3, x5); x1 = __hfma2(x1, x4, x6); x2 = __hfma2(x2, x5, x7); x3 = __hfma2(x3, x6, x0);
                    x4 = __hfma2(x4, x7, x1); x5 = __hfma2(x5, x0, x2); x6 = __hfma2(x6, x1, x3); x7 = __hfma2(x7, x2, x4);
                }
            }
            accum += __low2float(x0) + __high2float(x1) + __low2float(x2) + __high2float(x3) + __low2float(x4) + __high2float(x5) + __low2float(x6) + __high2float(x7);
        } else if (MODE == 4) {
            float x0 = ox, x1 = oy, x2 = oz, x3 = dx, x4 = dy, x5 = dz, x6 = ox + 1.f, x7 = oy + 1.f;
#pragma unroll 1
            for (int it = 0; it < NS; ++it) {
#pragma unroll
                for (int k = 0; k < 1; ++k) {
                    x0 = __fmaf_rn(x0, dx, dy); x1 = __fmaf_rn(x1, dx, dy); x2 = __fmaf_rn(x2, dx, dy); x3 = __fmaf_rn(x3, dx, dy);
                    x4 = __fmaf_rn(x4, dx, dz); x5 = __fmaf_rn(x5, dx, dz); x6 = __fmaf_rn(x6, dx, dz); x7 = __fmaf_rn(x7, dx, dz);
                }
            }
            accum += x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
        } else if (MODE == 5) {
            float2 x0 = make_float2(ox, oy), x1 = make_float2(oz, dx), x2 = make_float2(dy, dz), x3 = make_float2(ox + 1.f, oy + 1.f);
            const float2 m = make_float2(dx, dx), c = make_float2(dy, dz);
#pragma unroll 1
            for (int it = 0; it < NS; ++it) {
                x0 = __ffma2_rn(x0, m, c); x1 = __ffma2_rn(x1, m, c); x2 = __ffma2_rn(x2, m, c); x3 = __ffma2_rn(x3, m, c);
                x0 = __ffma2_rn(x0, m, c); x1 = __ffma2_rn(x1, m, c); x2 = __ffma2_rn(x2, m, c); x3 = __ffma2_rn(x3, m, c);
            }
            accum += x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y;
        }
        ox += 1e-3f; dz += 1e-4f; oy += 1e-3f; dy -= 1e-5f; oz += 1e-3f;
    }
    out[tid] = accum + ox;
    if (tid == 0) *clk = clock64() - t0;
    if (found) atomicAdd(cnt, found);
}

template <int MODE> void run(const char *name, const float4 *scene, float *out, unsigned *cnt, int sms, double mhz, double unit_per_scan) {
    static long long *clk = nullptr;
    if (!clk) CK(cudaMalloc(&clk, 8));
    const int grid = sms * 4;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    scan_kernel<MODE><<<grid, 256>>>(scene, out, cnt, clk);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(cnt, 0, 4));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 5; ++i) scan_kernel<MODE><<<grid, 256>>>(scene, out, cnt, clk);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= 5;
    unsigned c = 0;
    CK(cudaMemcpy(&c, cnt, 4, cudaMemcpyDeviceToHost));
    // warps per SMSP = 4 CTAs * 8 warps / 4 = 8; each does REPS scans of `unit_per_scan` units
    long long cy = 0;
    CK(cudaMemcpy(&cy, clk, 8, cudaMemcpyDeviceToHost));
    const double per_unit = ms * 1e-3 * mhz * 1e6 / (8.0 * REPS * unit_per_scan);
    printf("%-6s %8.3f ms  %lld clk (%.0f MHz eff)  %7.3f SMSP-clk per unit (unit/scan=%g)  found/scan-thread=%.3f\n", name, ms, cy, cy / (ms * 1e3), per_unit,
           unit_per_scan, (double)(c / 5) / ((double)grid * 256 * REPS));
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, %.0f MHz (attribute; compare with nvidia-smi under load)\n", p.name, p.multiProcessorCount, mhz);
    std::vector<float4> h(NS);
    srand(1);
    for (int i = 0; i < NS; ++i) {
        const float cx = -11.f + 22.f * rand() / RAND_MAX, cz = -11.f + 22.f * rand() / RAND_MAX;
        h[i] = make_float4(cx, 0.2f, cz, 0.2f);
    }
    std::vector<float4> hf(NS), hp(NS);
    for (int i = 0; i < NS; ++i) hf[i] = make_float4(h[i].x, h[i].y, h[i].z, -(h[i].x * h[i].x + h[i].y * h[i].y + h[i].z * h[i].z - h[i].w * h[i].w));
    for (int i = 0; i < NS; i += 2) {
        hp[i] = make_float4(hf[i].x, hf[i + 1].x, hf[i].y, hf[i + 1].y);
        hp[i + 1] = make_float4(hf[i].z, hf[i + 1].z, hf[i].w, hf[i + 1].w);
    }
    // half-precision records: hB = (cx_a,cx_b)(cy_a,cy_b)(cz_a,cz_b)(-r2_a,-r2_b) per slot pair; hA = (cx,cy)(cz,-r2) per slot
    std::vector<__half2> hb(2 * NS), ha(2 * NS);
    for (int i = 0; i < NS; i += 2) {
        hb[2 * i + 0] = __floats2half2_rn(h[i].x, h[i + 1].x); hb[2 * i + 1] = __floats2half2_rn(h[i].y, h[i + 1].y);
        hb[2 * i + 2] = __floats2half2_rn(h[i].z, h[i + 1].z); hb[2 * i + 3] = __floats2half2_rn(-h[i].w * h[i].w, -h[i + 1].w * h[i + 1].w);
    }
    for (int i = 0; i < NS; ++i) { ha[2 * i] = __floats2half2_rn(h[i].x, h[i].y); ha[2 * i + 1] = __floats2half2_rn(h[i].z, -h[i].w * h[i].w); }
    float4 *scene_hb, *scene_ha;
    CK(cudaMalloc(&scene_hb, sizeof(float4) * NS));
    CK(cudaMalloc(&scene_ha, sizeof(float4) * NS));
    CK(cudaMemset(scene_hb, 0, sizeof(float4) * NS));
    CK(cudaMemset(scene_ha, 0, sizeof(float4) * NS));
    CK(cudaMemcpy(scene_hb, hb.data(), sizeof(__half2) * 2 * NS, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(scene_ha, ha.data(), sizeof(__half2) * 2 * NS, cudaMemcpyHostToDevice));
    float4 *scene_f, *scene_p;
    CK(cudaMalloc(&scene_f, sizeof(float4) * NS));
    CK(cudaMalloc(&scene_p, sizeof(float4) * NS));
    CK(cudaMemcpy(scene_f, hf.data(), sizeof(float4) * NS, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(scene_p, hp.data(), sizeof(float4) * NS, cudaMemcpyHostToDevice));
    float4 *scene; float *out; unsigned *cnt;
    CK(cudaMalloc(&scene, sizeof(float4) * NS));
    CK(cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 4 * 256));
    CK(cudaMalloc(&cnt, 4));
    CK(cudaMemcpy(scene, h.data(), sizeof(float4) * NS, cudaMemcpyHostToDevice));
    const int sms = p.multiProcessorCount;
    run<0>("v14", scene, out, cnt, sms, mhz, NS);
    run<1>("v10", scene_f, out, cnt, sms, mhz, NS);
    run<2>("p8", scene_p, out, cnt, sms, mhz, NS);
    run<3>("p7", scene_p, out, cnt, sms, mhz, NS);
    run<6>("r2s16", scene_f, out, cnt, sms, mhz, NS);
    run<7>("r2p16", scene_f, out, cnt, sms, mhz, NS);
    run<8>("r2s32", scene_f, out, cnt, sms, mhz, NS);
    run<9>("r2p16c", scene_f, out, cnt, sms, mhz, NS);
    run<10>("hB", scene_hb, out, cnt, sms, mhz, NS);
    run<11>("hA", scene_ha, out, cnt, sms, mhz, NS);
    run<12>("hfma2", scene, out, cnt, sms, mhz, NS * 8.0);
    run<13>("hfma2v", scene, out, cnt, sms, mhz, NS * 8.0);
    run<4>("ffma", scene, out, cnt, sms, mhz, NS * 8.0);
    run<5>("ffma2", scene, out, cnt, sms, mhz, NS * 8.0);
    return 0;
}
