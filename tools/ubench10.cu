// The fill kernel's inner evaluation pipeline (row_prefetch + row_powf + fp32 tail, rolled row-pair loop) in isolation:
// no DP ring, no problem queue, no step bookkeeping.  Measures sub-partition cycles per predecessor evaluation for
// 1..4 warps per sub-partition.  The -DSA_KO_* knock-out switches it was run with lived in sa_kernels.cu up to commit 322adb9 (removed afterwards).
// Developer tool (round 1 cost model); includes the product kernels' source to reuse their device functions.
#include "../ampliconreconstructorom_b200/csrc/sa_kernels.cu"  // NOTE: written against the dense pipeline (SA_PRUNE=0 era); row_prefetch/row_powf still exist
using namespace sa;
template <int TAIL>
__global__ void __launch_bounds__(512) evalk(float *out, int steps, const float *dxs, FillParams P) {
    extern __shared__ unsigned char s_dyn[];
    PowTabSmem &s_tab = *tab_place(s_dyn);
    __shared__ float s_rowop[16][2 * LB][32];
    __shared__ float s_ring[16][LB][40];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    powtab_init(s_tab);
    for (int k = 0; k < LB; k++) { s_rowop[warp][k][lane] = 3000.0f * (k + 1) + 37.3f * lane + 0.7f * warp; s_rowop[warp][k + 6][lane] = 100.0f * k; }
    for (int k = 0; k < LB; k++) s_ring[warp][k][lane + 6] = 1000.0f * k + lane, s_ring[warp][k][lane & 7] = 5.0f;
    __syncthreads();
    const TabAddr a_tab = tab_addr(s_tab);
    const uint32_t a_row = (uint32_t)__cvta_generic_to_shared(&s_rowop[warp][0][lane]);
    const uint32_t a_my = (uint32_t)__cvta_generic_to_shared(&s_ring[warp][0][6 + lane]);
    uint32_t pc[LB];
    for (int d = 0; d < LB; d++) pc[d] = a_my + d * 160;
    float acc = 0.0f;
    for (int st = 0; st < steps; st++) {
        float dx[LB], fx[LB];
        const float4 a = __ldg(reinterpret_cast<const float4 *>(dxs) + (st & 63) * 3), b4 = __ldg(reinterpret_cast<const float4 *>(dxs) + (st & 63) * 3 + 1),
                     c4 = __ldg(reinterpret_cast<const float4 *>(dxs) + (st & 63) * 3 + 2);
        dx[0] = a.x, dx[1] = a.y, dx[2] = a.z, dx[3] = a.w, dx[4] = b4.x, dx[5] = b4.y;
        fx[0] = b4.z, fx[1] = b4.w, fx[2] = c4.x, fx[3] = c4.y, fx[4] = c4.z, fx[5] = c4.w;
        float best = -CUDART_INF_F;
        uint32_t umax = 0;
        PowRow rowA, rowB;
        row_prefetch(rowA, lds_f32(a_row), dx, a_tab, umax);
#ifdef UB_FULL_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int i = 0; i < LB / 2; i++) {
            const int dj = 2 * i + 1;
            row_prefetch(rowB, lds_f32(a_row + dj * 128), dx, a_tab, umax);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int djh = dj + h;
                float delta[LB], sc[LB];
#pragma unroll
                for (int c = 0; c < LB; c++) {
#ifdef SA_KO_PRED
                    sc[c] = __uint_as_float(pc[c] - djh * 4);
#else
                    sc[c] = lds_f32(pc[c] - djh * 4);
#endif
                }
                if (h == 0) {
                    row_powf(rowA, a_tab, P.pk, delta);
                    if (i < LB / 2 - 1) row_prefetch(rowA, lds_f32(a_row + (dj + 1) * 128), dx, a_tab, umax);
                } else row_powf(rowB, a_tab, P.pk, delta);
                if (TAIL) {
                    const float fn = 5000.0f * (float)(djh - 1);
#pragma unroll
                    for (int c = 0; c < LB; c++) {
                        const float fnfp = __fadd_rn(fn, fx[c]);
                        const float r = __fsub_rn(10000.0f, __fadd_rn(fnfp, delta[c]));
                        best = fmaxf(best, __fadd_rn(sc[c], r));
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < LB; c++) best = fmaxf(best, __fadd_rn(delta[c], sc[c]));
                }
            }
        }
        acc += (umax >= IX_RANGE) ? 1.0f : best;
        // rotate the column pointers like the kernel does
        const uint32_t t0 = pc[LB - 1];
#pragma unroll
        for (int d = LB - 1; d > 0; d--) pc[d] = pc[d - 1];
        pc[0] = t0;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount; const double ghz = p.clockRate * 1e-6;
    float *out, *dxs; cudaMalloc(&out, sms * 512 * 4); cudaMalloc(&dxs, 64 * 12 * 4);
    float h[64 * 12];
    for (int s = 0; s < 64; s++) for (int k = 0; k < 12; k++) h[s * 12 + k] = k < 6 ? 2500.0f * (k + 1) + 13.7f * s : 50.0f * (k - 6);
    cudaMemcpy(dxs, h, sizeof(h), cudaMemcpyHostToDevice);
    FillParams P; memset(&P, 0, sizeof(P)); fill_powf_constants(P.pk);
    cudaFuncSetAttribute(evalk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAB_SMEM_BYTES);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps : {1, 2, 3, 4}) {
        const int steps = 4000;
        evalk<1><<<sms, 128 * wps, TAB_SMEM_BYTES>>>(out, 50, dxs, P);
        cudaEventRecord(e0); evalk<1><<<sms, 128 * wps, TAB_SMEM_BYTES>>>(out, steps, dxs, P); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double clk = (ms * 1e-3 * ghz * 1e9) / ((double)steps * 36) / wps;
        printf("  w%d: %5.1f clk/eval (fp64 util %.2f)", wps, clk, 34.0 / clk);
    }
    printf("  %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
