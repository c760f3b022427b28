// Best-case throughput of the fill kernel's INSTRUCTION MIX with no dependency stalls: per "evaluation" 17 fp64
// (9 DFMA, 5 DMUL, 3 DADD) + NOTHER other instructions of the classes the kernel uses, 8 independent chains,
// 4 or 8 warps per SM sub-partition.  Tells how far the instruction mix itself is from the fp64 issue peak.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(double *sink, int iters, const double *cst) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x + 512] = 1.0f;
    __syncthreads();
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
    double a[8]; float f[8]; unsigned n[8];
    const double c0 = cst[0], c1 = cst[1];
    for (int u = 0; u < 8; u++) { a[u] = threadIdx.x * 1e-3 + u; f[u] = threadIdx.x + u; n[u] = threadIdx.x * 77 + u; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            // 17 fp64 on chain u (dependent inside the chain, 8 chains independent)
#pragma unroll
            for (int r = 0; r < 3; r++) { a[u] = __fma_rn(a[u], c0, c1); a[u] = __fma_rn(a[u], c1, c0); a[u] = __fma_rn(a[u], c0, c1); }
#pragma unroll
            for (int r = 0; r < 5; r++) a[u] = __dmul_rn(a[u], c0);
#pragma unroll
            for (int r = 0; r < 3; r++) a[u] = __dadd_rn(a[u], c1);
            if (MODE >= 1) {  // fp32 adds: 6
#pragma unroll
                for (int r = 0; r < 6; r++) f[u] = __fadd_rn(f[u], 1.25f);
            }
            if (MODE >= 2) {  // integer ALU: ~8 (LOP3, IADD3, SHF mixes)
#pragma unroll
                for (int r = 0; r < 4; r++) { n[u] = (n[u] >> 3) ^ (n[u] + 0x9e3779b9u); }
            }
            if (MODE >= 3) {  // 3 LDS (conflict-free, per-lane address)
                float v0, v1, v2;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(sbase + ((n[u] & 255u) << 2)));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(sbase + (((n[u] >> 8) & 255u) << 2)));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v2) : "r"(sbase + (((n[u] >> 16) & 255u) << 2)));
                f[u] = __fadd_rn(f[u], __fadd_rn(v0, __fadd_rn(v1, v2)));
            }
            if (MODE >= 4) {  // 2 conversions
                float g = __double2float_rn(a[u]);
                a[u] = __dadd_rn(a[u], (double)__int_as_float((__float_as_int(g) & 0x007fffff) | 0x3f800000));
            }
        }
    }
    double s = 0; for (int u = 0; u < 8; u++) s += a[u] + f[u] + n[u];
    if (s == 123.456) sink[0] = s;
}
template <int MODE> void run(const char *name, double *sink, const double *cst, int sms, int nother) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps : {4, 8}) {
        int threads = 128 * wps > 512 ? 512 : 128 * wps, ctas = (128 * wps) / threads, iters = 200;
        k<MODE><<<sms * ctas, threads>>>(sink, 5, cst);
        cudaEventRecord(e0); k<MODE><<<sms * ctas, threads>>>(sink, iters, cst); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double clk_per_eval = (ms * 1e-3 * 1.9e9) / (iters * 8.0) / wps;
        int nfp = (MODE >= 4) ? 18 : 17;
        printf("%-34s warps/SMSP=%d: %.1f clk per evaluation per SMSP  (fp64 pipe floor %d, issue floor %d) -> fp64 util %.2f\n",
               name, wps, clk_per_eval, 2 * nfp, nfp + nother, 2.0 * nfp / clk_per_eval);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink, *cst; cudaMalloc(&sink, 64); cudaMalloc(&cst, 64);
    double h[2] = {1.0000001, 1e-9}; cudaMemcpy(cst, h, 16, cudaMemcpyHostToDevice);
    int sms = p.multiProcessorCount;
    run<0>("fp64 only (17)", sink, cst, sms, 0);
    run<1>("+6 FADD", sink, cst, sms, 6);
    run<2>("+6 FADD +~12 int", sink, cst, sms, 18);
    run<3>("+6 FADD +12 int +3 LDS +3 FADD", sink, cst, sms, 24);
    run<4>("+ 2 conversions (+1 DADD, +2 int)", sink, cst, sms, 28);
    return 0;
}
