// In-silico digestion on the GPU: the motif scan of generate_cmap.py (/root/reference/generate_cmap.py:119-123),
// SURVEY.md section 8 row f-4 -- the producer of the CMAPs the DP path consumes.
//
//   indVPos = [m.start() + offset      for m in re.finditer('(?=' + motif + ')', seq)]        forward strand
//   indVNeg = [m.start() + rev_offset  for m in re.finditer('(?=' + revcomp(motif) + ')', seq)]  reverse strand
//   allPosns = sorted(set(indVNeg + indVPos))
//
// HBM-bound byte scan: 1 algorithmic byte per base.  Each thread owns 16 consecutive bases (one 16-byte vector
// load) plus an 8-byte look-ahead; upper-casing is one AND per 32-bit word; every position costs one funnel shift
// and two 32-bit compares (first four motif letters, both strands), the rest of the motif is only compared on a
// prefix hit (1 in 128 positions).  Matches are rare (one per ~5 kb), so they are appended through one global
// atomic counter and sorted + de-duplicated on the host (a few 10^5 integers per genome).
#include <algorithm>
#include <cstring>
#include <vector>
#include "sa_kernels.cuh"

namespace sa {

struct DigestParams {
    const uint8_t *seq;  // n bases, padded with >= 32 zero bytes, 16-byte aligned
    long long n;
    uint32_t f_lo, f_hi, f_mask_lo, f_mask_hi;  // forward motif, little-endian packed, with byte masks
    uint32_t r_lo, r_hi, r_mask_lo, r_mask_hi;  // reverse-complement motif
    int f_shift, r_shift;                        // label position = match start + shift
    long long *out;
    unsigned long long *count;
    long long cap;
};

__device__ __forceinline__ void emit(const DigestParams &P, long long pos) {
    const unsigned long long k = atomicAdd(P.count, 1ull);
    if ((long long)k < P.cap) P.out[k] = pos;
}

__global__ void __launch_bounds__(256) digest_kernel(const DigestParams P) {
    const long long n_vec = (P.n + 15) >> 4;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (long long)gridDim.x * blockDim.x) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(P.seq) + v);
        const uint2 b = __ldg(reinterpret_cast<const uint2 *>(P.seq + ((v + 1) << 4)));
        uint32_t w[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
#pragma unroll
        for (int k = 0; k < 6; k++) w[k] &= 0xDFDFDFDFu;  // str.upper() for ASCII letters (generate_cmap.py:106)
        const long long base = v << 4;
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const int k = p >> 2, o = (p & 3) * 8;
            const uint32_t lo = __funnelshift_r(w[k], w[k + 1], o);
            const bool fl = lo == P.f_lo, rl = lo == P.r_lo;  // motifs have >= 4 letters: the low word needs no mask
            if (fl || rl) {
                const uint32_t hi = __funnelshift_r(w[k + 1], (k + 2 < 6) ? w[k + 2] : 0u, o);
                const long long s = base + p;
                if (s < P.n) {
                    if (fl && (hi & P.f_mask_hi) == P.f_hi) emit(P, s + P.f_shift);
                    if (rl && (hi & P.r_mask_hi) == P.r_hi) emit(P, s + P.r_shift);
                }
            }
        }
    }
}

void launch_digest(const DigestParams &p, int n_sms, cudaStream_t st) {
    const long long n_vec = (p.n + 15) >> 4;
    const long long want = (n_vec + 255) / 256;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(want, (long long)n_sms * 16));
    digest_kernel<<<blocks, 256, 0, st>>>(p);
}

}  // namespace sa

// ---- C ABI ---------------------------------------------------------------------------------------------------------


namespace {
inline char comp(char c) {
    switch (c) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
        default: return 'N';
    }
}
void pack(const char *m, int len, uint32_t &lo, uint32_t &hi, uint32_t &mlo, uint32_t &mhi) {
    uint64_t bits = 0, mask = 0;
    for (int k = 0; k < len; k++) {
        bits |= (uint64_t)(uint8_t)m[k] << (8 * k);
        mask |= (uint64_t)0xFF << (8 * k);
    }
    lo = (uint32_t)bits;
    hi = (uint32_t)(bits >> 32);
    mlo = (uint32_t)mask;
    mhi = (uint32_t)(mask >> 32);
}
}  // namespace

// Implemented here (needs the kernel), called by sa_digest in sa_capi.cu which owns the context internals.
int sa_digest_run(cudaStream_t st, int n_sms, const uint8_t *d_seq, int64_t n, const char *motif, int32_t motif_len, int32_t offset,
                  long long *d_out, unsigned long long *d_count, int64_t cap) {
    if (motif_len < 4 || motif_len > 8) return SA_EINVAL;
    char fwd[8], rev[8];
    for (int k = 0; k < motif_len; k++) fwd[k] = (char)(motif[k] & 0xDF);
    for (int k = 0; k < motif_len; k++) rev[k] = comp(fwd[motif_len - 1 - k]);  // rev_complement (generate_cmap.py:38-39)
    sa::DigestParams P;
    memset(&P, 0, sizeof(P));
    P.seq = d_seq;
    P.n = n;
    pack(fwd, motif_len, P.f_lo, P.f_hi, P.f_mask_lo, P.f_mask_hi);
    pack(rev, motif_len, P.r_lo, P.r_hi, P.r_mask_lo, P.r_mask_hi);
    P.f_shift = offset;              // m.start() + offset          (offset already includes the "+1, cmap 1 based")
    P.r_shift = motif_len - offset;  // m.start() + rev_offset, rev_offset = len(regExpEnz) - offset
    P.out = d_out;
    P.count = d_count;
    P.cap = cap;
    sa::launch_digest(P, n_sms, st);
    return SA_OK;
}
