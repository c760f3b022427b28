// fp64 pipe capacity for the EXACT powf replay dataflow (17 fp64 per chain, 6 chains breadth-first), operands in
// registers, no table loads, no conversions (was built with -DSA_KO_EXP -DSA_KO_F2F_OUT, switches that lived in sa_kernels.cu up to commit 322adb9).  Developer tool.
#include "../ampliconreconstructorom_b200/csrc/sa_kernels.cu"
using namespace sa;
__global__ void __launch_bounds__(512) k11(float *out, int iters, FillParams P) {
    PowRow w;
    for (int c = 0; c < LB; c++) { w.z[c] = 0.8 + 0.01 * c + 1e-4 * threadIdx.x; w.invc[c] = 1.2 - 0.01 * c; w.y0[c] = 3.0 + c; }
    TabAddr ta; ta.t2d = 0; ta.exp = 0;
    float acc = 0;
    for (int it = 0; it < iters; it++) {
        float d[LB];
        row_powf(w, ta, P.pk, d);
#pragma unroll
        for (int c = 0; c < LB; c++) { acc += d[c]; w.z[c] = __hiloint2double(__double2hiint(w.z[c]), __double2loint(w.z[c]) ^ __float_as_int(d[c])); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount; const double ghz = p.clockRate * 1e-6;
    float *out; cudaMalloc(&out, sms * 512 * 4);
    FillParams P; memset(&P, 0, sizeof(P)); fill_powf_constants(P.pk);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps : {1, 2, 3, 4}) {
        const int iters = 20000;
        k11<<<sms, 128 * wps>>>(out, 50, P);
        cudaEventRecord(e0); k11<<<sms, 128 * wps>>>(out, iters, P); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double clk = (ms * 1e-3 * ghz * 1e9) / ((double)iters * 6) / wps;
        printf("  w%d: %5.1f clk/eval (fp64 util %.2f)", wps, clk, 34.0 / clk);
    }
    printf("  %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
