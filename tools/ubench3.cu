// Which instruction classes co-issue freely with DFMA? 6 DFMA chains + 9 "other" ops per round, 8 warps/SMSP.
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int ND, int NX>
__global__ void __launch_bounds__(1024) k(double *sink, int iters) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double a[8];
    int n[9]; float f[9];
    for (int u = 0; u < 8; u++) a[u] = threadIdx.x * 1e-3 + u;
    for (int u = 0; u < 9; u++) { n[u] = threadIdx.x * 7 + u; f[u] = threadIdx.x + u; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int u = 0; u < ND; u++) a[u] = __fma_rn(a[u], 1.0000001, 1e-9);
#pragma unroll
            for (int u = 0; u < NX; u++) {
                if (KIND == 0) n[u] = n[u] * 3 + 12345;                      // IMAD
                if (KIND == 1) n[u] = (n[u] ^ 0x5bd1e995) & (n[u] + 0);      // LOP3 (may fold) 
                if (KIND == 2) f[u] = __fadd_rn(f[u], 1.5f);                 // FADD
                if (KIND == 3) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f[u]) : "r"((unsigned)(((n[u] >> 2) & 1023) * 4)));  // LDS
                if (KIND == 4) n[u] = __vimax3_s32(n[u], n[(u + 1) % 9], 77);  // VIMNMX3 (ALU)
                if (KIND == 5) n[u] = n[u] + 0x1234567;                        // IADD (ALU)
                if (KIND == 6) f[u] = __fmul_rn(f[u], 1.0001f);               // FMUL
                if (KIND == 7) n[u] = (n[u] >> 3) ^ n[u];                     // SHF + LOP
            }
        }
    }
    double s = 0; for (int u = 0; u < 8; u++) s += a[u];
    float m = 0; for (int u = 0; u < 9; u++) m += n[u] + f[u];
    if (s == 123.456 || m == 12345.f) sink[0] = s + m;
}
template <int KIND, int ND, int NX> void run(const char *name, double *sink, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 300;
    k<KIND, ND, NX><<<sms, 1024>>>(sink, 10);
    cudaEventRecord(e0); k<KIND, ND, NX><<<sms, 1024>>>(sink, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-10s ND=%d NX=%d: %.2f clk per round per warp-slot (8 warps/SMSP); ideal issue=%d, fp64 pipe=%d\n", name, ND, NX,
           (ms * 1e-3 * 1.9e9) / (iters * 16.0) / 8.0, ND + NX, 2 * ND);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink; cudaMalloc(&sink, 64);
    int sms = p.multiProcessorCount;
    run<0, 6, 0>("none", sink, sms);
    run<0, 6, 9>("IMAD", sink, sms); run<1, 6, 9>("LOP3", sink, sms); run<2, 6, 9>("FADD", sink, sms); run<3, 6, 9>("LDS", sink, sms);
    run<4, 6, 9>("VIMNMX3", sink, sms); run<5, 6, 9>("IADD", sink, sms); run<6, 6, 9>("FMUL", sink, sms); run<7, 6, 9>("SHF+LOP", sink, sms);
    run<0, 0, 9>("IMAD only", sink, sms); run<2, 0, 9>("FADD only", sink, sms); run<5, 0, 9>("IADD only", sink, sms); run<3, 0, 9>("LDS only", sink, sms);
    run<0, 6, 6>("IMAD", sink, sms); run<5, 6, 6>("IADD", sink, sms); run<0, 6, 3>("IMAD", sink, sms); run<5, 6, 3>("IADD", sink, sms);
    return 0;
}
