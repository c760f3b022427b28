// Does the fp64 pipe sustain its peak when every DFMA reads three DISTINCT register pairs (as in the powf replay),
// rather than one register + two loop-invariant operands (as in the peak microbenchmark)?  Developer tool.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(double *sink, int iters, const double *cst) {
    constexpr int NCH = 6;
    double x[NCH], y[NCH], z[NCH];
    const double c0 = cst[0], c1 = cst[1];
    for (int u = 0; u < NCH; u++) { x[u] = threadIdx.x * 1e-3 + u; y[u] = 1.0 + u * 1e-9; z[u] = 1e-9 * u; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 6; s++) {
#pragma unroll
            for (int u = 0; u < NCH; u++) {
                if (MODE == 0) { x[u] = __fma_rn(x[u], c0, c1); y[u] = __fma_rn(y[u], c0, c1); z[u] = __fma_rn(z[u], c0, c1); }   // 1 reg + 2 invariant
                if (MODE == 1) { x[u] = __fma_rn(x[u], y[u], c1); y[u] = __fma_rn(y[u], z[u], c1); z[u] = __fma_rn(z[u], x[u], c1); }  // 2 regs
                if (MODE == 2) { x[u] = __fma_rn(x[u], y[u], z[u]); y[u] = __fma_rn(y[u], z[u], x[u]); z[u] = __fma_rn(z[u], x[u], y[u]); }  // 3 regs
                if (MODE == 3) { x[u] = __dmul_rn(x[u], y[u]); y[u] = __dadd_rn(y[u], z[u]); z[u] = __dmul_rn(z[u], x[u]); }  // 2-reg mul/add
                if (MODE == 4) { double t = __fma_rn(x[u], y[u], z[u]); double w = __fma_rn(y[u], z[u], x[u]); z[u] = __fma_rn(z[u], x[u], y[u]); x[u] = t; y[u] = w; }  // 3 regs, new dest
            }
        }
    }
    double s = 0;
    for (int u = 0; u < NCH; u++) s += x[u] + y[u] + z[u];
    if (s == 123.456) sink[0] = s;
}
template <int MODE> void run(const char *name, double *sink, const double *cst, int sms, double ghz) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%-36s", name);
    for (int wps : {1, 2, 3, 4}) {
        int threads = 128 * wps, iters = 500;
        k<MODE><<<sms, threads>>>(sink, 5, cst);
        cudaEventRecord(e0); k<MODE><<<sms, threads>>>(sink, iters, cst); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double clk = (ms * 1e-3 * ghz * 1e9) / ((double)iters * 6 * 6 * 3) / wps;  // cycles per fp64 warp-instruction per SMSP
        printf("  w%d: %.2f clk/instr (util %.2f)", wps, clk, 2.0 / clk);
    }
    printf("\n");
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink, *cst; cudaMalloc(&sink, 64); cudaMalloc(&cst, 64);
    double h[2] = {1.0000001, 1e-9}; cudaMemcpy(cst, h, 16, cudaMemcpyHostToDevice);
    int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
    run<0>("DFMA r, inv, inv", sink, cst, sms, ghz);
    run<1>("DFMA r, r, inv", sink, cst, sms, ghz);
    run<2>("DFMA r, r, r (in place)", sink, cst, sms, ghz);
    run<4>("DFMA r, r, r (new dest)", sink, cst, sms, ghz);
    run<3>("DMUL/DADD r, r", sink, cst, sms, ghz);
    return 0;
}
