// Instruction-throughput microbenchmarks for the pipes the fill kernel leans on (developer probe).
// Reports thread-level instructions per clock per SM for each op, measured with all SMs busy.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int OP>
__global__ void __launch_bounds__(256) k(double *sink, int iters) {
    __shared__ double sm[512];
    for (int i = threadIdx.x; i < 512; i += 256) sm[i] = i * 1.0;
    __syncthreads();
    double a[8];
    float f[8];
    int n[8];
    for (int u = 0; u < 8; u++) { a[u] = threadIdx.x * 1e-3 + u; f[u] = threadIdx.x * 1e-3f + u; n[u] = threadIdx.x + u; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (OP == 0) a[u] = __fma_rn(a[u], 1.0000001, 1e-9);
                if (OP == 1) a[u] = __dadd_rn(a[u], 1e-9);
                if (OP == 2) a[u] = __dmul_rn(a[u], 1.0000001);
                if (OP == 3) { a[u] = (double)f[u]; f[u] = __int_as_float(__double2hiint(a[u]) ^ 0x3f800001); }   // F2F.F64.F32 (+ LOP)
                if (OP == 4) { f[u] = __double2float_rn(a[u]); a[u] = __hiloint2double(__float_as_int(f[u]) | 0x3ff00000, n[u]); }  // F2F.F32.F64 (+LOP)
                if (OP == 5) { a[u] = (double)n[u]; n[u] = __double2hiint(a[u]) ^ n[u]; }  // I2F.F64 (+LOP)
                if (OP == 6) { n[u] = (n[u] * 1103515245 + 12345); a[u] = sm[(n[u] >> 16) & 511]; n[u] ^= __double2loint(a[u]); }  // LDS.64 random
                if (OP == 7) { n[u] = n[u] * 5 + 1; f[u] = __int_as_float(n[u]) ; }  // IMAD baseline
                if (OP == 8) { f[u] = __shfl_sync(0xffffffffu, f[u], (n[u] + u) & 31); }  // SHFL
            }
    }
    double s = 0; for (int u = 0; u < 8; u++) s += a[u] + f[u] + n[u];
    if (s == 123.456) sink[0] = s;
}
template <int OP> void run(const char *name, double *sink, int sms, double extra) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 2000, ctas = sms * 8;
    k<OP><<<ctas, 256>>>(sink, 10);
    cudaEventRecord(e0); k<OP><<<ctas, 256>>>(sink, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)ctas * 256 * iters * 64;
    printf("%-22s %8.3f ms  %.3e thread-inst/s  = %.1f /clk/SM @1.9GHz %s\n", name, ms, inst / (ms * 1e-3), inst / (ms * 1e-3) / sms / 1.9e9, extra ? "(+1 int op each)" : "");
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink; cudaMalloc(&sink, 64);
    int sms = p.multiProcessorCount;
    run<0>("DFMA", sink, sms, 0); run<1>("DADD", sink, sms, 0); run<2>("DMUL", sink, sms, 0);
    run<3>("F2F.F64.F32", sink, sms, 1); run<4>("F2F.F32.F64", sink, sms, 1); run<5>("I2F.F64.S32", sink, sms, 1);
    run<6>("LDS.64 random", sink, sms, 1); run<7>("IMAD", sink, sms, 0); run<8>("SHFL", sink, sms, 0);
    return 0;
}
