// DFMA latency/throughput vs resident warps and independent chains; and DFMA mixed with ALU work (developer probe).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH, int MIX>
__global__ void k(double *sink, int iters) {
    double a[CH];
    int n[8];
    for (int u = 0; u < CH; u++) a[u] = threadIdx.x * 1e-3 + u;
    for (int u = 0; u < 8; u++) n[u] = threadIdx.x + u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int u = 0; u < CH; u++) a[u] = __fma_rn(a[u], 1.0000001, 1e-9);
            if (MIX) {
#pragma unroll
                for (int u = 0; u < (CH * MIX) / 2; u++) n[u % 8] = n[u % 8] * 3 + n[(u + 1) % 8];   // IMAD per iteration: MIX/2 per DFMA
            }
        }
    }
    double s = 0; for (int u = 0; u < CH; u++) s += a[u];
    int m = 0; for (int u = 0; u < 8; u++) m += n[u];
    if (s == 123.456 || m == 12345) sink[0] = s + m;
}
template <int CH, int MIX> void run(double *sink, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps : {1, 2, 4, 8, 16}) {   // warps per SMSP
        int threads = 128 * wps > 1024 ? 1024 : 128 * wps; int ctas_per_sm = (128 * wps) / threads;
        int iters = 400;
        k<CH, MIX><<<sms * ctas_per_sm, threads>>>(sink, 10);
        cudaEventRecord(e0); k<CH, MIX><<<sms * ctas_per_sm, threads>>>(sink, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double inst = (double)sms * 128 * wps * iters * 16 * CH;
        printf("chains=%d mix=%d warps/SMSP=%2d : %6.1f DFMA/clk/SM (@1.9GHz)  cyc per DFMA-round per warp = %.1f\n", CH, MIX, wps,
               inst / (ms * 1e-3) / sms / 1.9e9, (ms * 1e-3 * 1.9e9) / (iters * 16.0));
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink; cudaMalloc(&sink, 64);
    int sms = p.multiProcessorCount;
    run<1, 0>(sink, sms); run<2, 0>(sink, sms); run<4, 0>(sink, sms); run<6, 0>(sink, sms); run<8, 0>(sink, sms);
    run<6, 3>(sink, sms);   // 1.5 IMAD per DFMA, like the fill kernel's 28:18
    return 0;
}
