// Cost of one memory-class instruction per "evaluation" (17 fp64 on 6 chains + ALU/FADD filler) on B200:
// LDS.32 / LDS.64 / LDS.128 (conflict free), SHFL, STS.32, LDG (L1 hit).  Developer tool (round 1 cost model).
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int NM>
__global__ void __launch_bounds__(512) k(double *sink, int iters, const double *cst, const float *g) {
    __shared__ __align__(16) float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const unsigned sb32 = (unsigned)__cvta_generic_to_shared(sm) + lane * 4;
    const unsigned sb64 = (unsigned)__cvta_generic_to_shared(sm) + (lane & 15) * 8;
    const unsigned sb128 = (unsigned)__cvta_generic_to_shared(sm) + (lane & 7) * 16;
    constexpr int NCH = 6;
    double a[NCH]; float f[NCH]; unsigned n[NCH];
    const double c0 = cst[0], c1 = cst[1];
    for (int u = 0; u < NCH; u++) { a[u] = threadIdx.x * 1e-3 + u; f[u] = threadIdx.x + u; n[u] = threadIdx.x * 77 + u; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 17; s++) {
#pragma unroll
            for (int u = 0; u < NCH; u++) {
                if (s % 3 == 0) a[u] = __fma_rn(a[u], c0, c1);
                else if (s % 3 == 1) a[u] = __dmul_rn(a[u], c0);
                else a[u] = __dadd_rn(a[u], c1);
                if (s < 4) n[u] = (n[u] ^ 0x9e3779b9u) + (n[u] >> 3);
                if (s < 6) f[u] = __fadd_rn(f[u], 1.25f);
                if (s < NM) {
                    const unsigned r = (n[u] & 7u) << 7;
                    if (KIND == 0) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sb32 + r)); f[u] = __fadd_rn(f[u], v); }
                    if (KIND == 1) { float v, w; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v), "=f"(w) : "r"(sb64 + r)); f[u] = __fadd_rn(f[u], __fadd_rn(v, w)); }
                    if (KIND == 2) { float v, w, x, y; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v), "=f"(w), "=f"(x), "=f"(y) : "r"(sb128 + r)); f[u] = __fadd_rn(f[u], __fadd_rn(__fadd_rn(v, w), __fadd_rn(x, y))); }
                    if (KIND == 3) { float v = __shfl_sync(0xffffffffu, f[u], (int)(n[u] & 31u)); f[u] = __fadd_rn(f[u], v); }
                    if (KIND == 4) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(sb32 + r), "f"(f[u]) : "memory"); }
                    if (KIND == 5) { float v = __ldg(g + lane + (n[u] & 7u) * 32); f[u] = __fadd_rn(f[u], v); }
                    if (KIND == 6) { float v = __shfl_up_sync(0xffffffffu, f[u], 1); f[u] = __fadd_rn(f[u], v); }
                }
            }
        }
    }
    double s = 0;
    for (int u = 0; u < NCH; u++) s += a[u] + f[u] + n[u];
    if (s == 123.456) sink[0] = s;
}
template <int KIND, int NM> void run(const char *name, double *sink, const double *cst, const float *g, int sms, double ghz) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%-28s", name);
    for (int wps : {1, 3, 4}) {
        int threads = 128 * wps, iters = 300;
        k<KIND, NM><<<sms, threads>>>(sink, 5, cst, g);
        cudaEventRecord(e0); k<KIND, NM><<<sms, threads>>>(sink, iters, cst, g); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double clk = (ms * 1e-3 * ghz * 1e9) / ((double)iters * 6) / wps;
        printf("  w%d: %5.1f clk/eval", wps, clk);
    }
    printf("\n");
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink, *cst; float *g; cudaMalloc(&sink, 64); cudaMalloc(&cst, 64); cudaMalloc(&g, 4096); cudaMemset(g, 0, 4096);
    double h[2] = {1.0000001, 1e-9}; cudaMemcpy(cst, h, 16, cudaMemcpyHostToDevice);
    int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
    printf("cycles per evaluation per SM sub-partition (17 fp64 => floor 34); N = memory-class instructions per evaluation\n");
    run<0, 0>("none", sink, cst, g, sms, ghz);
    run<0, 1>("LDS.32 x1", sink, cst, g, sms, ghz);
    run<0, 3>("LDS.32 x3", sink, cst, g, sms, ghz);
    run<1, 1>("LDS.64 x1", sink, cst, g, sms, ghz);
    run<1, 3>("LDS.64 x3", sink, cst, g, sms, ghz);
    run<2, 1>("LDS.128 x1", sink, cst, g, sms, ghz);
    run<2, 3>("LDS.128 x3", sink, cst, g, sms, ghz);
    run<3, 1>("SHFL.IDX x1", sink, cst, g, sms, ghz);
    run<3, 3>("SHFL.IDX x3", sink, cst, g, sms, ghz);
    run<6, 3>("SHFL.UP x3", sink, cst, g, sms, ghz);
    run<4, 1>("STS.32 x1", sink, cst, g, sms, ghz);
    run<4, 3>("STS.32 x3", sink, cst, g, sms, ghz);
    run<5, 1>("LDG(L1) x1", sink, cst, g, sms, ghz);
    run<5, 3>("LDG(L1) x3", sink, cst, g, sms, ghz);
    return 0;
}
