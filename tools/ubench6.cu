// How does the fp64 pipe of one SM sub-partition behave when DFMA/DMUL/DADD are mixed with the other instruction
// classes of the fill kernel?  Per "evaluation": 17 fp64 on NCH independent chains + NA integer ALU + NF FADD +
// NL LDS.32 (conflict free) + NX conversions (F2F.F64.F32 + F2F.F32.F64 pairs count 2).  Reports fp64 pipe
// utilisation for 1..6 warps per sub-partition.  Developer tool (round 1 roofline model), not part of the product.
#include <cstdio>
#include <cuda_runtime.h>
template <int NCH, int NA, int NF, int NL, int NX>
__global__ void __launch_bounds__(768) k(double *sink, int iters, const double *cst) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4;
    double a[NCH];
    float f[NCH];
    unsigned n[NCH];
    const double c0 = cst[0], c1 = cst[1];
    for (int u = 0; u < NCH; u++) { a[u] = threadIdx.x * 1e-3 + u; f[u] = threadIdx.x + u; n[u] = threadIdx.x * 77 + u; }
    for (int it = 0; it < iters; it++) {
        // breadth-first: stage s of all chains, then stage s+1
#pragma unroll
        for (int s = 0; s < 17; s++) {
#pragma unroll
            for (int u = 0; u < NCH; u++) {
                if (s % 3 == 0) a[u] = __fma_rn(a[u], c0, c1);
                else if (s % 3 == 1) a[u] = __dmul_rn(a[u], c0);
                else a[u] = __dadd_rn(a[u], c1);
                // sprinkle the other instructions evenly over the 17 stages
                if (s < NA) n[u] = (n[u] ^ 0x9e3779b9u) + (n[u] >> 3);   // 2-3 ALU ops
                if (s < NF) f[u] = __fadd_rn(f[u], 1.25f);
                if (s < NL) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sbase + ((n[u] & 7u) << 7))); f[u] = __fadd_rn(f[u], v); }
                if (s < NX) { float g = __double2float_rn(a[u]); n[u] ^= __float_as_uint(g); }
            }
        }
    }
    double s = 0;
    for (int u = 0; u < NCH; u++) s += a[u] + f[u] + n[u];
    if (s == 123.456) sink[0] = s;
}
template <int NCH, int NA, int NF, int NL, int NX> void run(const char *name, double *sink, const double *cst, int sms, double ghz) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%-44s", name);
    for (int wps : {1, 2, 3, 4, 6}) {
        int threads = 128 * wps, iters = 300;
        k<NCH, NA, NF, NL, NX><<<sms, threads>>>(sink, 5, cst);
        cudaEventRecord(e0); k<NCH, NA, NF, NL, NX><<<sms, threads>>>(sink, iters, cst); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double clk_per_eval = (ms * 1e-3 * ghz * 1e9) / ((double)iters * NCH) / wps;  // per SMSP
        printf("  w%d: %5.1f clk (util %.2f)", wps, clk_per_eval, 34.0 / clk_per_eval);
    }
    printf("\n");
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink, *cst; cudaMalloc(&sink, 64); cudaMalloc(&cst, 64);
    double h[2] = {1.0000001, 1e-9}; cudaMemcpy(cst, h, 16, cudaMemcpyHostToDevice);
    int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
    printf("SMs %d clock %.3f GHz; clk = cycles per 17-fp64 evaluation per SM sub-partition (floor 34)\n", sms, ghz);
    run<6, 0, 0, 0, 0>("6ch fp64 only", sink, cst, sms, ghz);
    run<6, 4, 0, 0, 0>("6ch +4x(2-3 ALU)", sink, cst, sms, ghz);
    run<6, 8, 0, 0, 0>("6ch +8x(2-3 ALU)", sink, cst, sms, ghz);
    run<6, 4, 6, 0, 0>("6ch +4xALU +6 FADD", sink, cst, sms, ghz);
    run<6, 4, 6, 3, 0>("6ch +4xALU +6 FADD +3 LDS(+3 FADD)", sink, cst, sms, ghz);
    run<6, 4, 6, 0, 2>("6ch +4xALU +6 FADD +2 F2F", sink, cst, sms, ghz);
    run<6, 4, 6, 3, 2>("6ch +4xALU +6 FADD +3 LDS +2 F2F", sink, cst, sms, ghz);
    run<3, 4, 6, 3, 2>("3ch +4xALU +6 FADD +3 LDS +2 F2F", sink, cst, sms, ghz);
    run<12, 4, 6, 3, 2>("12ch +4xALU +6 FADD +3 LDS +2 F2F", sink, cst, sms, ghz);
    return 0;
}
