// In-silico digestion on the GPU: the motif scan of generate_cmap.py (/root/reference/generate_cmap.py:119-123),
// SURVEY.md section 8 row f-4 -- the producer of the CMAPs the DP path consumes.
//
//   indVPos = [m.start() + offset      for m in re.finditer('(?=' + motif + ')', seq)]        forward strand
//   indVNeg = [m.start() + rev_offset  for m in re.finditer('(?=' + revcomp(motif) + ')', seq)]  reverse strand
//   allPosns = sorted(set(indVNeg + indVPos))
//
// HBM-bound byte scan: 1 algorithmic byte per base.  Each thread owns 32 consecutive bases per iteration (two 16-byte vector
// loads) plus an 8-byte look-ahead, and has the next iteration's loads in flight while it works on the current ones.  Bits 1..2 of an ASCII base are a 2-bit code that does not depend on case (A/a 00, C/c 01, T/t 10,
// G/g 11), so the 32 + 8 bases are first PACKED 2 bits each: per 32-bit word one AND and one multiply gather the four codes
// into the top byte, byte permutes assemble one 32-bit window per 16 bases (11 instructions each).  A motif letter k is then
// tested at all 16 positions at once: funnel-shift the window by 2k bits, XOR with the letter's code replicated 16 times,
// OR into a mismatch word (one shift shared by both strands + one LOP3 per strand and letter); a position whose 2-bit
// field stays zero is a candidate.  ~3 thread-instructions per base instead of one shift, two compares and a branch per
// base AND strand, and no divergent branch in the common case.  Other characters (N, IUPAC codes, the zero padding) alias
// to some base code, so every candidate is verified against the real bytes (case-folded like str.upper()) before it is
// emitted; candidates are rare (one per ~8 kb per strand) and go through one global atomic counter, and the host sorts
// and de-duplicates them (a few 10^5 integers per genome).
#include <algorithm>
#include <cstring>
#include <vector>
#include "sa_kernels.cuh"

namespace sa {

struct DigestParams {
    const uint8_t *seq;  // n bases, padded with >= 32 zero bytes, 16-byte aligned
    long long n;
    uint32_t fm[8], rm[8];  // per letter: 2-bit code x 0x55555555, forward motif / reverse-complement motif
    uint8_t fl[8], rl[8];   // the letters themselves (upper case) for the verification
    int len;                // motif length, 4..8
    int f_shift, r_shift;   // label position = match start + shift
    long long *out;
    unsigned long long *count;
    long long cap;
};

__device__ __forceinline__ void emit(const DigestParams &P, long long pos) {
    const unsigned long long k = atomicAdd(P.count, 1ull);
    if ((long long)k < P.cap) P.out[k] = pos;
}

// four bases (one 32-bit word) -> their four 2-bit codes in the TOP byte (base j of the word at bits 24 + 2j)
__device__ __forceinline__ uint32_t pack4(uint32_t w) {
    // fields at bits 1-2, 9-10, 17-18, 25-26 move to 24-25, 26-27, 28-29, 30-31; every cross term lands below bit 24 or
    // beyond bit 31, and no two terms overlap (no carries)
    return (w & 0x06060606u) * ((1u << 23) + (1u << 17) + (1u << 11) + (1u << 5));
}

// rare path: candidates of one strand (bit 2p set = position p of this thread's 16), verified byte by byte
__device__ __noinline__ void digest_verify(const DigestParams &P, long long base, uint32_t cand, const uint8_t *letters, int shift) {
    while (cand) {
        const int p = (__ffs(cand) - 1) >> 1;
        cand &= cand - 1;
        const long long s = base + p;
        if (s >= P.n) continue;
        bool ok = true;
        for (int k = 0; k < P.len; k++) ok = ok && ((__ldg(P.seq + s + k) & 0xDFu) == letters[k]);  // str.upper() (generate_cmap.py:106)
        if (ok) emit(P, s + shift);
    }
}

// mismatch words of one 16-base window (w0 = its 2-bit codes, w1 = the codes that follow): a 2-bit field of tf / tr stays
// zero where every letter of the forward / reverse-complement motif matches
template <int LEN>
__device__ __forceinline__ void digest_match(const DigestParams &P, uint32_t w0, uint32_t w1, uint32_t &tf, uint32_t &tr) {
    tf = w0 ^ P.fm[0];
    tr = w0 ^ P.rm[0];
#pragma unroll
    for (int k = 1; k < LEN; k++) {
        const uint32_t x = __funnelshift_r(w0, w1, 2 * k);
        tf |= x ^ P.fm[k];
        tr |= x ^ P.rm[k];
    }
}
// nonzero iff some 2-bit field of t is zero (borrows can only add false alarms above a real zero field)
__device__ __forceinline__ uint32_t has_zero_field(uint32_t t) { return (t - 0x55555555u) & ~t & 0xAAAAAAAAu; }
__device__ __forceinline__ uint32_t zero_fields(uint32_t t) { return ~(t | (t >> 1)) & 0x55555555u; }  // bit 2p = field p is zero

struct DigestChunk {  // 32 bases + 8 bases of look-ahead
    uint4 a, b;
    uint2 l;
};
__device__ __forceinline__ DigestChunk digest_load(const DigestParams &P, int v) {
    DigestChunk c;
    const uint4 *src = reinterpret_cast<const uint4 *>(P.seq) + 2ll * v;
    c.a = __ldg(src);
    c.b = __ldg(src + 1);
    c.l = __ldg(reinterpret_cast<const uint2 *>(src + 2));
    return c;
}

template <int LEN>
__global__ void __launch_bounds__(256) digest_kernel(const __grid_constant__ DigestParams P) {
    const int n_chunk = (int)((P.n + 31) >> 5);  // < 2^31: sa_digest_run rejects sequences of 2^36 bases and more
    const int stride = (int)(gridDim.x * blockDim.x);
    int v = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (v >= n_chunk) return;
    DigestChunk cur = digest_load(P, v);
    for (; v < n_chunk; v += stride) {
        const int vn = (v + stride < n_chunk) ? v + stride : v;  // software pipelining: the next chunk is in flight while this one is matched
        const DigestChunk nxt = digest_load(P, vn);
        // 2-bit windows: base p of the first / second 16 at bits 2p..2p+1 of wa / wb, the 8 look-ahead bases in the low half of wl
        const uint32_t wa = __byte_perm(__byte_perm(pack4(cur.a.x), pack4(cur.a.y), 0x0073), __byte_perm(pack4(cur.a.z), pack4(cur.a.w), 0x0073), 0x5410);
        const uint32_t wb = __byte_perm(__byte_perm(pack4(cur.b.x), pack4(cur.b.y), 0x0073), __byte_perm(pack4(cur.b.z), pack4(cur.b.w), 0x0073), 0x5410);
        const uint32_t wl = __byte_perm(pack4(cur.l.x), pack4(cur.l.y), 0x0073);
        uint32_t tfa, tra, tfb, trb;
        digest_match<LEN>(P, wa, wb, tfa, tra);
        digest_match<LEN>(P, wb, wl, tfb, trb);
        if ((has_zero_field(tfa) | has_zero_field(tra)) | (has_zero_field(tfb) | has_zero_field(trb))) {
            const long long base = (long long)v << 5;
            const uint32_t cfa = zero_fields(tfa), cra = zero_fields(tra), cfb = zero_fields(tfb), crb = zero_fields(trb);
            if (cfa) digest_verify(P, base, cfa, P.fl, P.f_shift);
            if (cra) digest_verify(P, base, cra, P.rl, P.r_shift);
            if (cfb) digest_verify(P, base + 16, cfb, P.fl, P.f_shift);
            if (crb) digest_verify(P, base + 16, crb, P.rl, P.r_shift);
        }
        cur = nxt;
    }
}

template <int LEN>
static void launch_digest_len(const DigestParams &p, int n_sms, cudaStream_t st) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, digest_kernel<LEN>, 256, 0);
    const long long n_chunk = (p.n + 31) >> 5;
    const long long want = (n_chunk + 255) / 256;
    // one resident wave: every CTA streams its share from start to end (no second wave, no tail of late CTAs)
    const int blocks = (int)std::max<long long>(1, std::min<long long>(want, (long long)n_sms * std::max(per_sm, 1)));
    digest_kernel<LEN><<<blocks, 256, 0, st>>>(p);
}

void launch_digest(const DigestParams &p, int n_sms, cudaStream_t st) {
    switch (p.len) {
        case 4: launch_digest_len<4>(p, n_sms, st); break;
        case 5: launch_digest_len<5>(p, n_sms, st); break;
        case 6: launch_digest_len<6>(p, n_sms, st); break;
        case 7: launch_digest_len<7>(p, n_sms, st); break;
        default: launch_digest_len<8>(p, n_sms, st); break;
    }
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
}  // namespace

// Implemented here (needs the kernel), called by sa_digest in sa_capi.cu which owns the context internals.
int sa_digest_run(cudaStream_t st, int n_sms, const uint8_t *d_seq, int64_t n, const char *motif, int32_t motif_len, int32_t offset,
                  long long *d_out, unsigned long long *d_count, int64_t cap) {
    if (motif_len < 4 || motif_len > 8 || n >= (1ll << 36)) return SA_EINVAL;
    char fwd[8], rev[8];
    for (int k = 0; k < motif_len; k++) fwd[k] = (char)(motif[k] & 0xDF);
    for (int k = 0; k < motif_len; k++) rev[k] = comp(fwd[motif_len - 1 - k]);  // rev_complement (generate_cmap.py:38-39)
    sa::DigestParams P;
    memset(&P, 0, sizeof(P));
    P.seq = d_seq;
    P.n = n;
    for (int k = 0; k < motif_len; k++) {
        if (!fwd[k] || !strchr("ACGT", fwd[k])) return SA_EINVAL;  // the enzyme table holds ACGT motifs only (generate_cmap.py:60-68)
        P.fl[k] = (uint8_t)fwd[k];
        P.rl[k] = (uint8_t)rev[k];
        P.fm[k] = (((uint32_t)(uint8_t)fwd[k] >> 1) & 3u) * 0x55555555u;
        P.rm[k] = (((uint32_t)(uint8_t)rev[k] >> 1) & 3u) * 0x55555555u;
    }
    P.len = motif_len;
    P.f_shift = offset;              // m.start() + offset          (offset already includes the "+1, cmap 1 based")
    P.r_shift = motif_len - offset;  // m.start() + rev_offset, rev_offset = len(regExpEnz) - offset
    P.out = d_out;
    P.count = d_count;
    P.cap = cap;
    sa::launch_digest(P, n_sms, st);
    return SA_OK;
}
