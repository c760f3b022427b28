// Issue-rate microbenchmarks for the bound pass of the fill kernel on sm_100a: packed fp32 (FADD2/FFMA2), 3-input
// FMNMX3, MUFU.LG2/EX2, LDS.32/LDS.64, and the per-candidate-pair instruction mixes (current scalar form vs packed form).
// Prints warp-instructions per clock per SM sub-partition (4 warps resident per sub-partition).  Developer tool.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N_ACC 8
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ float lo2(u64 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi2(u64 v) { return __uint_as_float((uint32_t)(v >> 32)); }

template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, int iters) {
    __shared__ __align__(16) float sh[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = 1.0f + i * 1e-3f;
    __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sh) + (threadIdx.x & 31) * 8;
    float a[N_ACC];
    u64 p[N_ACC];
#pragma unroll
    for (int i = 0; i < N_ACC; i++) { a[i] = 1.0f + threadIdx.x * 1e-4f + i; p[i] = pk2(a[i], a[i] + 0.5f); }
    const float c1 = 1.0000001f, c2 = 0.999f;
    const u64 pc = pk2(c1, c2);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N_ACC; i++) {
            if (MODE == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));
            if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
            if (MODE == 2) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));
            if (MODE == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
            if (MODE == 4) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 6) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sbase + i * 256)); a[i] += v; }
            if (MODE == 7) { u64 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(sbase + i * 256)); a[i] += lo2(v) + hi2(v); }
            if (MODE == 8) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pc));
            if (MODE == 9) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < N_ACC; i++) s += a[i] + lo2(p[i]) + hi2(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Per-pair mixes of the bound pass.  MIX 0: scalar form (what fill_kernel does today, 2 candidates):
//   2 LDS, 2 FSUB, 2 LG2, 2 FFMA, 2 EX2, 2 FADD(fnfp), 2 FADD, 2 FSUB, 2 FADD, + top-2 tracking (10 instr)
// MIX 1: packed form: 1 LDS.64, 1 FADD2, 2 LG2, 1 FFMA2, 2 EX2, 1 FADD2.RP, 1 FADD2.RP + the same tracking
// MIX 2: packed form with a cheaper tracking (per 6 candidates: 2 FMNMX3 + 3 for second max ... modelled as 4 per pair)
// MIX 3: MIX 1 without MUFU (issue-only cost)   MIX 4: tracking only
template <int MIX>
__global__ void __launch_bounds__(256, 3) kmix(float *out, int iters, float seed) {
    __shared__ __align__(16) float sh[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 1000.0f + i * 0.37f;
    __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sh) + (threadIdx.x & 31) * 8 + (threadIdx.x >> 5) * 1024;
    float m1 = -1e30f, m2 = -1e30f;
    int k6 = 0;
    float db = seed + threadIdx.x, kc = 9000.0f + seed;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int pr = 0; pr < 18; pr++) {
            float h0, h1;
            const float dx0 = seed * (pr + 1), dx1 = seed * (pr + 2) + it;
            if (MIX == 0) {
                float s0, s1;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(s0) : "r"(sbase + (pr & 7) * 32));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(s1) : "r"(sbase + (pr & 7) * 32 + 4));
                float d0 = __fsub_rn(db, dx0), d1 = __fsub_rn(db, dx1), l0, l1, e0, e1;
                asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(fabsf(d0)));
                asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(fabsf(d1)));
                l0 = __fmaf_rn(l0, 1.2f, -1e-5f), l1 = __fmaf_rn(l1, 1.2f, -1e-5f);
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(l0));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(l1));
                const float f0 = __fadd_rn(5000.0f, kc), f1 = __fadd_rn(5000.0f, dx1);
                h0 = __fadd_rn(s0, __fsub_rn(10000.0f, __fadd_rn(f0, e0)));
                h1 = __fadd_rn(s1, __fsub_rn(10000.0f, __fadd_rn(f1, e1)));
            } else if (MIX == 4) {
                h0 = dx0 + db, h1 = dx1 - db;
            } else {
                u64 s, d, l, e, h;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(s) : "r"(sbase + (pr & 7) * 32));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(db, db)), "l"(pk2(dx0, dx1)));
                float l0, l1, e0, e1;
                if (MIX != 3) {
                    asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(fabsf(lo2(d))));
                    asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(fabsf(hi2(d))));
                } else l0 = lo2(d), l1 = hi2(d);
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(l) : "l"(pk2(l0, l1)), "l"(pk2(1.2f, 1.2f)), "l"(pk2(-1e-5f, -1e-5f)));
                if (MIX != 3) {
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(lo2(l)));
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(hi2(l)));
                } else e0 = lo2(l), e1 = hi2(l);
                e = pk2(e0, e1);
                asm volatile("add.rp.f32x2 %0, %1, %2;" : "=l"(h) : "l"(s), "l"(pk2(kc, dx1)));
                asm volatile("sub.rp.f32x2 %0, %1, %2;" : "=l"(h) : "l"(h), "l"(e));
                h0 = lo2(h), h1 = hi2(h);
            }
            if (MIX == 2) {
                const float mx = fmaxf(h0, h1);
                float t;
                asm volatile("min.f32 %0, %1, %2;" : "=f"(t) : "f"(m1), "f"(mx));
                asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(m2) : "f"(m2), "f"(t), "f"(fminf(h0, h1)));
                k6 = (mx > m1) ? pr : k6;
                m1 = fmaxf(m1, mx);
            } else {
                const float mx = fmaxf(h0, h1), mn = fminf(h0, h1);
                const int kmx = (h0 >= h1) ? 2 * pr : 2 * pr + 1;
                const float t = fminf(m1, mx);
                k6 = (mx > m1) ? kmx : k6;
                m1 = fmaxf(m1, mx);
                m2 = fmaxf(m2, fmaxf(mn, t));
            }
        }
        db += m1 * 1e-9f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = m1 + m2 + k6;
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    float *out; cudaMalloc(&out, sms * 3 * 512 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[] = {"FADD", "FADD2", "FMNMX", "FMNMX3", "MUFU.LG2", "MUFU.EX2", "LDS.32(+FADD)", "LDS.64(+2FADD)", "FFMA2", "FFMA"};
    const int iters = 20000;
#define RUN(M) { k<M><<<sms, 512>>>(out, 100); cudaEventRecord(e0); k<M><<<sms, 512>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); \
        float ms; cudaEventElapsedTime(&ms, e0, e1); const double clk = ms * 1e-3 * ghz * 1e9; \
        printf("%-16s %6.3f clk per warp-instr per sub-partition (4 warps/SMSP)\n", names[M], clk / ((double)iters * N_ACC * 4)); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    const char *mix[] = {"scalar pair (today)", "packed pair", "packed pair + FMNMX3 tracking", "packed pair, no MUFU", "tracking only"};
#define RUNM(M) { kmix<M><<<sms * 3, 256>>>(out, 50, 1.5f); cudaEventRecord(e0); kmix<M><<<sms * 3, 256>>>(out, 4000, 1.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); \
        float ms; cudaEventElapsedTime(&ms, e0, e1); const double clk = ms * 1e-3 * ghz * 1e9; \
        printf("mix %-32s %7.1f clk per 36 candidates per sub-partition (6 warps/SMSP)\n", mix[M], clk / (4000.0 * 6)); }
    RUNM(0) RUNM(1) RUNM(2) RUNM(3) RUNM(4)
    printf("%s, %.3f GHz, %d SMs\n", cudaGetErrorString(cudaGetLastError()), ghz, sms);
    return 0;
}
