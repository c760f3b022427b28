// Register-file read bandwidth model: fp64 instructions interleaved with ALU instructions that read 1, 2 or 3
// DISTINCT registers.  If the SM sub-partition can read only ~2 x 32-bit register operands per cycle, the time per
// evaluation follows the operand-word count, not the instruction count.  Developer tool (round 1 cost model).
#include <cstdio>
#include <cuda_runtime.h>
template <int FMODE, int NALU, int AREGS>
__global__ void __launch_bounds__(512) k(double *sink, int iters, const double *cst, const unsigned *ui) {
    constexpr int NCH = 6;
    double x[NCH], y[NCH];
    unsigned n[NCH], m[NCH], q[NCH];
    const double c0 = cst[0], c1 = cst[1];
    for (int u = 0; u < NCH; u++) { x[u] = threadIdx.x * 1e-3 + u; y[u] = 1.0 + u * 1e-9; n[u] = ui[u] + threadIdx.x; m[u] = ui[u + 8] ^ threadIdx.x; q[u] = ui[u + 16] * threadIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 17; s++) {
#pragma unroll
            for (int u = 0; u < NCH; u++) {
                if (FMODE == 0) x[u] = __dadd_rn(x[u], 3.0);                 // r, imm      : 2 words
                if (FMODE == 1) { x[u] = __dadd_rn(x[u], y[u]); }            // r, r        : 4 words
                if (FMODE == 2) { x[u] = __fma_rn(x[u], y[u], c1); }         // r, r, inv   : 4 words (+reuse)
                if (s < NALU) {
                    if (AREGS == 1) n[u] = (n[u] >> 3) + 0x9e3779b9u;        // 1 register operand
                    if (AREGS == 2) n[u] = n[u] + m[u];                      // 2
                    if (AREGS == 3) { n[u] = n[u] + m[u] + q[u]; m[u] ^= 0x55u; }  // IADD3 r,r,r (+ a 1-operand op)
                }
            }
        }
    }
    double s = 0;
    for (int u = 0; u < NCH; u++) s += x[u] + y[u] + n[u] + m[u] + q[u];
    if (s == 123.456) sink[0] = s;
}
template <int FMODE, int NALU, int AREGS> void run(const char *name, double *sink, const double *cst, const unsigned *ui, int sms, double ghz) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%-44s", name);
    for (int wps : {1, 3, 4}) {
        int threads = 128 * wps, iters = 300;
        k<FMODE, NALU, AREGS><<<sms, threads>>>(sink, 5, cst, ui);
        cudaEventRecord(e0); k<FMODE, NALU, AREGS><<<sms, threads>>>(sink, iters, cst, ui); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("  w%d: %5.1f clk/eval", wps, (ms * 1e-3 * ghz * 1e9) / ((double)iters * 6) / wps);
    }
    printf("\n");
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink, *cst; unsigned *ui; cudaMalloc(&sink, 64); cudaMalloc(&cst, 64); cudaMalloc(&ui, 256); cudaMemset(ui, 1, 256);
    double h[2] = {1.0000001, 1e-9}; cudaMemcpy(cst, h, 16, cudaMemcpyHostToDevice);
    int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
    printf("17 fp64 per evaluation (pipe floor 34 clk per SM sub-partition) + NALU*? ALU instructions\n");
    run<0, 0, 1>("DADD r,imm", sink, cst, ui, sms, ghz);
    run<1, 0, 1>("DADD r,r", sink, cst, ui, sms, ghz);
    run<2, 0, 1>("DFMA r,r,inv", sink, cst, ui, sms, ghz);
    run<0, 12, 1>("DADD r,imm + 24 ALU(1 reg)", sink, cst, ui, sms, ghz);
    run<0, 17, 2>("DADD r,imm + 17 ALU(2 regs)", sink, cst, ui, sms, ghz);
    run<0, 12, 3>("DADD r,imm + 12 IADD3(3 regs) + 12 ALU(1)", sink, cst, ui, sms, ghz);
    run<1, 12, 1>("DADD r,r + 24 ALU(1 reg)", sink, cst, ui, sms, ghz);
    run<1, 17, 2>("DADD r,r + 17 ALU(2 regs)", sink, cst, ui, sms, ghz);
    run<1, 12, 3>("DADD r,r + 12 IADD3(3 regs) + 12 ALU(1)", sink, cst, ui, sms, ghz);
    run<2, 17, 2>("DFMA r,r,inv + 17 ALU(2 regs)", sink, cst, ui, sms, ghz);
    run<2, 12, 3>("DFMA r,r,inv + 12 IADD3(3 regs) + 12 ALU(1)", sink, cst, ui, sms, ghz);
    return 0;
}
