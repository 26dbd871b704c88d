// ubench_pipe.cu -- sm_100a FP32 pipe rates: FFMA vs FFMA2 (packed / broadcast operands) and mixes
// with LDS / SHF, at 8 warps per SMSP.  Reports SMSP clocks per warp-instruction of the named kind.
// build: nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -o ubench_pipe ubench_pipe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;
constexpr int BODY = 8;     // repetitions of the 8-wide group per loop iteration

__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(256, 4) pipe_kernel(float *out, float s0, float s1) {
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(s0, s1, s0, s1);
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = 1.f + 1e-3f * (tid + i);
    const float m = s0, c = s1;
    uint32_t sg = 0;
    uint32_t w[4] = {(uint32_t)tid, (uint32_t)tid * 3u, (uint32_t)tid * 5u, (uint32_t)tid * 7u};
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < BODY; ++r) {
            if (MODE == 0) {            // 16 scalar FFMA, constants m, c
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], m, c);
            } else if (MODE == 1) {     // 8 FFMA2, broadcast scalars
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float2 v = __ffma2_rn(make_float2(x[2 * i], x[2 * i + 1]), make_float2(m, m), make_float2(c, c));
                    x[2 * i] = v.x; x[2 * i + 1] = v.y;
                }
            } else if (MODE == 2) {     // 8 FFMA2, packed constants (m, c as distinct pairs)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float2 v = __ffma2_rn(make_float2(x[2 * i], x[2 * i + 1]), make_float2(m, c), make_float2(c, m));
                    x[2 * i] = v.x; x[2 * i + 1] = v.y;
                }
            } else if (MODE == 3) {     // 8 FFMA2 whose addend is another accumulator pair (3 packed register operands)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = (i + 3) & 7;
                    float2 v = __ffma2_rn(make_float2(x[2 * i], x[2 * i + 1]), make_float2(m, m), make_float2(x[2 * j], x[2 * j + 1]));
                    x[2 * i] = v.x; x[2 * i + 1] = v.y;
                }
            } else if (MODE == 4) {     // 16 FFMA with 3 varying register operands
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], x[(i + 5) & 15], x[(i + 11) & 15]);
            } else if (MODE == 5) {     // 16 FFMA + 2 LDS.128 + 2 SHF (the v10 mix without dependence on the loads)
                const float4 a = lds4(base + ((it + r) & 15) * 16), b = lds4(base + 256 + ((it + r) & 15) * 16);
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], m, (i & 1) ? a.x : b.y);
                sg = __funnelshift_l(__float_as_uint(x[0]), sg, 1);
                sg = __funnelshift_l(__float_as_uint(x[1]), sg, 1);
            } else if (MODE == 6) {     // 8 FFMA2 (broadcast) + 2 LDS.128 + 2 SHF
                const float4 a = lds4(base + ((it + r) & 15) * 16), b = lds4(base + 256 + ((it + r) & 15) * 16);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float2 v = __ffma2_rn(make_float2(x[2 * i], x[2 * i + 1]), make_float2(m, m), (i & 1) ? make_float2(a.x, a.y) : make_float2(b.z, b.w));
                    x[2 * i] = v.x; x[2 * i + 1] = v.y;
                }
                sg = __funnelshift_l(__float_as_uint(x[0]), sg, 1);
                sg = __funnelshift_l(__float_as_uint(x[1]), sg, 1);
            } else if (MODE >= 10 && MODE < 20) {   // 8 FFMA2 (broadcast) + K funnel shifts on independent words
                constexpr int K = (MODE - 10) * 2;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float2 v = __ffma2_rn(make_float2(x[2 * i], x[2 * i + 1]), make_float2(m, m), make_float2(c, c));
                    x[2 * i] = v.x; x[2 * i + 1] = v.y;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) w[k & 3] = __funnelshift_l(w[(k + 1) & 3], w[k & 3], 1);
            } else if (MODE >= 20 && MODE < 30) {   // 8 FFMA2 + K LDS.128 (results folded in rarely)
                constexpr int K = (MODE - 20) * 2;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float2 v = __ffma2_rn(make_float2(x[2 * i], x[2 * i + 1]), make_float2(m, m), make_float2(c, c));
                    x[2 * i] = v.x; x[2 * i + 1] = v.y;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) { const float4 a = lds4(base + ((it + r * K + k) & 63) * 16); w[k & 3] ^= __float_as_uint(a.x) ^ __float_as_uint(a.w); }
            } else if (MODE >= 30 && MODE < 40) {   // 16 FFMA (const) + K funnel shifts
                constexpr int K = (MODE - 30) * 2;
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], m, c);
#pragma unroll
                for (int k = 0; k < K; ++k) w[k & 3] = __funnelshift_l(w[(k + 1) & 3], w[k & 3], 1);
            } else if (MODE == 7) {     // 8 FFMA + 4 FFMA2: does the packed form add FP32 throughput next to scalar FFMA?
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __fmaf_rn(x[i], m, c);
#pragma unroll
                for (int i = 4; i < 8; ++i) {
                    float2 v = __ffma2_rn(make_float2(x[2 * i], x[2 * i + 1]), make_float2(m, m), make_float2(c, c));
                    x[2 * i] = v.x; x[2 * i + 1] = v.y;
                }
            }
        }
    }
    float acc = (float)(sg + w[0] + w[1] + w[2] + w[3]);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += x[i];
    out[tid] = acc;
}

template <int MODE> void run(const char *name, float *out, int sms, double mhz, double fma_lane_ops_per_group, double instr_per_group) {
    const int grid = sms * 4;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    pipe_kernel<MODE><<<grid, 256>>>(out, 0.999f, 1e-3f);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 3; ++i) pipe_kernel<MODE><<<grid, 256>>>(out, 0.999f, 1e-3f);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= 3;
    const double clk = ms * 1e-3 * mhz * 1e6;                         // SM clocks of one launch (assuming mhz)
    const double groups = 8.0 * ITERS * BODY;                         // per SMSP: 8 warps
    printf("%-28s %8.3f ms  %6.2f clk per group | %5.3f clk per FMA-lane-op-per-lane | %5.3f clk per issued instr\n", name, ms,
           clk / groups, clk / (groups * fma_lane_ops_per_group), clk / (groups * instr_per_group));
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double mhz = khz / 1000.0;
    float *out;
    CK(cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 4 * 256));
    const int sms = p.multiProcessorCount;
    printf("%s %d SMs, assuming %.0f MHz\n", p.name, sms, mhz);
    run<0>("16 FFMA (const m,c)", out, sms, mhz, 16, 16);
    run<1>("8 FFMA2 broadcast", out, sms, mhz, 16, 8);
    run<2>("8 FFMA2 packed consts", out, sms, mhz, 16, 8);
    run<3>("8 FFMA2 3 packed regs", out, sms, mhz, 16, 8);
    run<4>("16 FFMA 3 varying regs", out, sms, mhz, 16, 16);
    run<5>("16 FFMA + 2 LDS + 2 SHF", out, sms, mhz, 16, 20);
    run<6>("8 FFMA2 + 2 LDS + 2 SHF", out, sms, mhz, 16, 12);
    run<7>("8 FFMA + 4 FFMA2", out, sms, mhz, 16, 12);
    run<11>("8 FFMA2 + 2 SHF", out, sms, mhz, 16, 10);
    run<12>("8 FFMA2 + 4 SHF", out, sms, mhz, 16, 12);
    run<14>("8 FFMA2 + 8 SHF", out, sms, mhz, 16, 16);
    run<16>("8 FFMA2 + 12 SHF", out, sms, mhz, 16, 20);
    run<21>("8 FFMA2 + 2 LDS(+2 LOP3)", out, sms, mhz, 16, 12);
    run<22>("8 FFMA2 + 4 LDS(+4 LOP3)", out, sms, mhz, 16, 16);
    run<31>("16 FFMA + 2 SHF", out, sms, mhz, 16, 18);
    run<32>("16 FFMA + 4 SHF", out, sms, mhz, 16, 20);
    run<34>("16 FFMA + 8 SHF", out, sms, mhz, 16, 24);
    return 0;
}
