// Does an XU conversion (F2F/I2F with 64-bit types) cost more than one issue slot?  Mixes per round:
// ND DFMA (independent chains) + NC conversions + NF FADDs, 4 warps/SMSP (like the fill kernel) and 8.
#include <cstdio>
#include <cuda_runtime.h>
template <int ND, int NC, int NF>
__global__ void k(double *sink, int iters) {
    double a[8]; float f[8]; double c[4]; float g[4];
    for (int u = 0; u < 8; u++) { a[u] = threadIdx.x * 1e-3 + u; f[u] = threadIdx.x + u; }
    for (int u = 0; u < 4; u++) { c[u] = threadIdx.x + 0.5 * u; g[u] = u; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int u = 0; u < ND; u++) a[u] = __fma_rn(a[u], 1.0000001, 1e-9);
#pragma unroll
            for (int u = 0; u < NC; u++) {   // alternate F2F.F32.F64 and F2F.F64.F32 through a dependent pair
                if (u & 1) c[u >> 1] = (double)g[u >> 1] + 0.0;   // F2F.F64.F32 (the +0.0 folds)
                else g[u >> 1] = __double2float_rn(c[u >> 1]);
            }
#pragma unroll
            for (int u = 0; u < NF; u++) f[u % 8] = __fadd_rn(f[u % 8], 1.5f);
        }
    }
    double s = 0; for (int u = 0; u < 8; u++) s += a[u] + f[u];
    for (int u = 0; u < 4; u++) s += c[u] + g[u];
    if (s == 123.456) sink[0] = s;
}
template <int ND, int NC, int NF> void run(double *sink, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps : {4, 8}) {
        int threads = 128 * wps, iters = 300;
        k<ND, NC, NF><<<sms, threads>>>(sink, 10);
        cudaEventRecord(e0); k<ND, NC, NF><<<sms, threads>>>(sink, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA=%d CONV=%d FADD=%d warps/SMSP=%d: %.2f clk per round per warp (issue=%d, fp64=%d, xu=%d)\n", ND, NC, NF, wps,
               (ms * 1e-3 * 1.9e9) / (iters * 16.0) / wps, ND + NC + NF, 2 * ND, 8 * NC);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink; cudaMalloc(&sink, 64);
    int sms = p.multiProcessorCount;
    run<6, 0, 10>(sink, sms); run<6, 1, 9>(sink, sms); run<6, 2, 8>(sink, sms); run<6, 4, 6>(sink, sms);
    run<0, 2, 14>(sink, sms); run<0, 4, 12>(sink, sms); run<0, 0, 16>(sink, sms);
    return 0;
}
