// sm_100a kernels of the SegAligner DP hot path.
//
//   fill_kernel    dp_align (/root/reference/SegAligner/SegAligner.cpp:95-133) with score_f (:83-92) inlined,
//                  the three initialisers (:254-287) folded in as the cell's initial value, and the
//                  backtrack-start scans (:142-154, :172-195) fused into the sweep.
//   trace_kernel   get_aln_list (SAHelper.cpp:83-95); recomputes the scores along the path when only the
//                  predecessor codes were stored.
//   detect_kernel  multi_backtrack_scores + zero_out_aln (SegAligner.cpp:198-233).
//
// Mapping.  One warp owns one (contig x segment) problem at a time (persistent warps pull problems, largest
// first, from a global queue; large problems are swept by a whole CTA or a thread-block cluster as a wavefront).
// Lanes lie along the contig (rows j) in strips of 32 rows; the warp sweeps the segment columns q = 0..nx-1 of a
// strip, then moves to the next strip.  Cell (j,q) depends only on rows j-6..j-1 and columns q-6..q-1
// (SegAligner.cpp:100,108-110), so all 32 cells of a step are independent.
//   * the last 6 columns of the strip live in a per-warp shared-memory ring, mirrored ([12][6+32]) so that the six
//     previous columns sit at static offsets; entries 0..5 of each slot are the 6 rows above the strip (bottom rows
//     of the previous strip, spilled to a small L2-resident scratch), so a predecessor read is one LDS with no
//     boundary special-casing: out-of-matrix predecessors simply read -inf;
//   * per-label operand tables (LabelOps) make `b[j]-b[i]`, `x[q]-x[p]` and the fp_t term single loads; the
//     row-side values sit in registers for the whole strip, the column-side values are warp-uniform loads;
//   * EXACT PRUNING (default): every candidate gets an fp32 upper bound of its score (two MUFU), only the winner's
//     powf(discrep,1.2f) is replayed bit-exactly in fp64 (powf12.cuh, tables in shared memory), and a cell whose
//     winner is not strictly ahead of the runner-up's bound is rescanned exactly (see "exact pruning" below).
// All fp32 arithmetic of the exact path uses explicit round-to-nearest intrinsics (no FMA contraction), in the
// reference's order.  Tie-breaking: the reference scans predecessors with i ascending then p ascending and replaces on
// strict '>' (:109-127), i.e. the winner is the max score with the smallest (i,p) = the largest (dj,dq); the exact
// rescan (cell_slow) scans dj, dq ascending and keeps the last maximum, which selects the same predecessor.
#include <cstdio>
#include <math_constants.h>
#include <cooperative_groups.h>
#include <cub/device/device_scan.cuh>
namespace cg = cooperative_groups;
#include "sa_kernels.cuh"
#include "powf12.cuh"

namespace sa {

__constant__ PowLogEntry c_logtab[16] = {SA_POWF_LOG2_TAB};
__constant__ unsigned long long c_exptab[32] = {SA_EXP2F_TAB};
// fp64 constants of the powf replay, read as constant-bank operands of DFMA/DMUL (no materialisation cost)
__constant__ double c_pk[12] = {SA_POWF_A0, SA_POWF_A1, SA_POWF_A2, SA_POWF_A3, SA_POWF_A4, SA_Y12,
                                SA_EXP2F_C0, SA_EXP2F_C1, SA_EXP2F_C2, SA_EXP2F_SHIFT, -SA_EXP2F_SHIFT, 0.0};

// ---- shared-memory tables of the powf replay -----------------------------------------------------------------
// Variant knob (compile time, both measured on B200 -- see DESIGN.md 4.1):
//   SA_TAB2D 3 (dense kernel, SA_PRUNE=0): LANE-REPLICATED, bank-conflict-free tables.  The log table {invc_i, logc_i + (double)k}
//                 holds 32 binades x 16 entries, each entry stored 8 times (16 B x 8 = one 128-byte bank row), and lane
//                 l reads copy l & 7: the 8 lanes of an LDS.128 quarter-warp phase always hit 8 different 16-byte bank
//                 groups => exactly 4 shared-memory wavefronts per LDS.128 whatever the indices.  The exp2 table holds
//                 its 32 u64 entries 16 times (8 B x 16 = one bank row), lane l reads copy l & 15 => exactly 2
//                 wavefronts per LDS.64.  64 KB + 4 KB per CTA, so the kernels run one CTA per SM.
//   SA_TAB2D 1 (pruning kernel, the default): one copy: 52 x 16 entries of 16 B + 32 x u64 (13.4 wavefronts per evaluation instead of 7: random
//                 indices conflict 2-3 ways per phase); 16 KB, 3 CTAs per SM.

// fast-path domain of the delta term: discrep in [0.7 * 2^-KLO, 0.7 * 2^(KBIN-KLO)); anything else (exact 0, tiny,
// huge, inf, NaN) is flagged and the cell is redone through the generic path
#if SA_TAB2D == 3
constexpr int KLO = 9, KBIN = 32;  // 1.4e-3 <= discrep < 5.9e6 bp
#else
constexpr int KLO = 20, KBIN = 52;
#endif
constexpr uint32_t IX_LO = SA_POWF_OFF - ((uint32_t)KLO << 23);
constexpr uint32_t IX_RANGE = (uint32_t)KBIN << 23;
constexpr int TAB2D_N = KBIN * 16;  // entry (k + KLO) * 16 + i

#define NEG_INF_F (__int_as_float(0xff800000))
__device__ __noinline__ float powf12_slow(float ax) {
    return powf12_generic(ax, c_logtab, reinterpret_cast<const uint64_t *>(c_exptab));
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void lds_f64x2(uint32_t addr, double &a, double &b) {
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void lds_u32x2(uint32_t addr, uint32_t &lo, uint32_t &hi) {
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr));
}

#if SA_TAB2D == 3
constexpr int T2D_COPIES = 8, EXP_COPIES = 16;
struct __align__(16) PowTabSmem {  // placed so that t2d sits at a 64 KB-aligned SHARED address (tab_place)
    unsigned long long exp64[32 * EXP_COPIES];  // 4 KB, [entry][copy]
    double2 t2d[TAB2D_N * T2D_COPIES];          // 64 KB, [entry][copy] = {invc, y0}
};
static_assert(sizeof(PowTabSmem) == 4096 + 65536, "PowTabSmem layout");
constexpr uint32_t TAB_ALIGN = 65536, TAB_LEAD = 4096;
constexpr uint32_t T2D_SHIFT = 12, T2D_MASK = 0xFF80u;  // byte offset of entry u >> 19: (u >> 12) & 0xFF80
constexpr uint32_t EXP_SHIFT = 7, EXP_MASK = 0xF80u;    // byte offset of entry ki & 31: (ki << 7) & 0xF80
#else
struct __align__(16) PowTabSmem {  // placed at a 16 KB-aligned SHARED address by the kernels (tab_place)
    double2 t2d[TAB2D_N];           // {invc, y0}; byte offset = ((u >> 15) & 0x3FF0) stays inside this struct
    unsigned long long exp64[32];   // at byte 13312
    unsigned char pad[16384 - TAB2D_N * 16 - 256];
};
static_assert(sizeof(PowTabSmem) == 16384, "PowTabSmem layout");
constexpr uint32_t TAB_ALIGN = 16384, TAB_LEAD = 0;
constexpr uint32_t T2D_SHIFT = 15, T2D_MASK = 0x3FF0u;
constexpr uint32_t EXP_SHIFT = 3, EXP_MASK = 0xF8u;
#endif

// The `base | offset` addressing needs the log table at a shared-window address aligned to its (power-of-two) size;
// static shared variables start at a toolchain-chosen offset, so kernels over-allocate dynamic shared memory and
// align here.  TAB_SMEM_BYTES is what a kernel must reserve (the struct lands somewhere inside it).
constexpr uint32_t TAB_SMEM_BYTES = (uint32_t)sizeof(PowTabSmem) + TAB_ALIGN;
__device__ __forceinline__ PowTabSmem *tab_place(unsigned char *dyn) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(dyn);
    const uint32_t a = (base + TAB_LEAD + TAB_ALIGN - 1) & ~(TAB_ALIGN - 1);  // shared address of t2d
    return reinterpret_cast<PowTabSmem *>(dyn + (a - TAB_LEAD - base));
}

// Exact fp32 <-> fp64 conversions with integer ALU instructions.  On B200 the XU conversions (F2F) share an issue
// path with the fp64 pipe and the LSU -- measured (tools/ubench6.cu): every F2F costs ~5.5 and every LDS ~8.4
// sub-partition cycles ON TOP of the fp64 time, while ALU/FMA-pipe instructions overlap with it for free.
#ifndef SA_INT_CVT
#define SA_INT_CVT 1
#endif
// widen a NORMAL fp32 bit pattern (here: iz in [0x3f330000, 0x3fb30000)) to fp64
__device__ __forceinline__ double f2d_normal(uint32_t fbits) {
#if SA_INT_CVT
    return __hiloint2double((int)((fbits >> 3) + 0x38000000u), (int)(fbits << 29));
#else
    return (double)__uint_as_float(fbits);
#endif
}
// round a positive fp64 whose value lies in the NORMAL fp32 range to fp32, round-to-nearest-even (== cvt.rn.f32.f64)
__device__ __forceinline__ float d2f_normal(double y) {
#if SA_INT_CVT
    const uint32_t hi = (uint32_t)__double2hiint(y), lo = (uint32_t)__double2loint(y);
    const uint32_t shf = __funnelshift_l(lo, hi, 3);  // drops sign + 2 exponent bits, keeps 29 + 3 mantissa bits
    const uint32_t rem = lo & 0x1FFFFFFFu;            // the 29 bits below the fp32 mantissa
    const uint32_t inc = (rem + 0x0FFFFFFFu + (shf & 1u)) >> 29;  // 1 iff rem > half, or == half and lsb odd
    return __uint_as_float(shf + 0x40000000u + inc);  // + 0x80 on the 9 kept exponent bits == exponent - 896
#else
    return __double2float_rn(y);
#endif
}

// Per-lane base addresses (32-bit shared window) of the two tables, lane copy folded in.
struct TabAddr {
    uint32_t t2d, exp;
};
__device__ __forceinline__ TabAddr tab_addr(const PowTabSmem &t) {
    TabAddr a;
    const uint32_t lane = threadIdx.x & 31;
    a.t2d = (uint32_t)__cvta_generic_to_shared(&t.t2d[0]);
    a.exp = (uint32_t)__cvta_generic_to_shared(&t.exp64[0]);
#if SA_TAB2D == 3
    a.t2d |= (lane & (T2D_COPIES - 1)) << 4;
    a.exp |= (lane & (EXP_COPIES - 1)) << 3;
#else
    (void)lane;
#endif
    return a;
}

__device__ __forceinline__ void powtab_init(PowTabSmem &t) {
#if SA_TAB2D == 3
    for (int e = threadIdx.x; e < TAB2D_N * T2D_COPIES; e += blockDim.x) {
        const int ent = e / T2D_COPIES, k = (ent >> 4) - KLO, i = ent & 15;
        t.t2d[e] = make_double2(c_logtab[i].invc, __dadd_rn(c_logtab[i].logc, (double)k));
    }
    for (int e = threadIdx.x; e < 32 * EXP_COPIES; e += blockDim.x) t.exp64[e] = c_exptab[e / EXP_COPIES];
#else
    for (int e = threadIdx.x; e < TAB2D_N; e += blockDim.x) {
        const int k = (e >> 4) - KLO, i = e & 15;
        t.t2d[e] = make_double2(c_logtab[i].invc, __dadd_rn(c_logtab[i].logc, (double)k));
    }
    if (threadIdx.x < 32) t.exp64[threadIdx.x] = c_exptab[threadIdx.x];
#endif
}

// Branch-free replay of powf(ax, 1.2f) for ax inside the fast domain; outside it the result is garbage but the
// code is still memory-safe (every table index is masked), so callers evaluate unconditionally and patch the
// rare out-of-domain lanes afterwards (u_out >= IX_RANGE tells which).  ta = per-lane table addresses (tab_addr);
// the tables are aligned to their size, so `base | offset` == `base + offset` (one LOP3 per address).
__device__ __forceinline__ float powf12_fast(float ax, const TabAddr ta, const double (&pk)[12], uint32_t &u_out) {
    const uint32_t ix = __float_as_uint(ax);
    const uint32_t u = ix - IX_LO;
    u_out = u;
    const uint32_t kb = u & 0xff800000u;
    const uint32_t iz = ix - kb + ((uint32_t)KLO << 23);
    double invc, y0;
    lds_f64x2(ta.t2d | ((u >> T2D_SHIFT) & T2D_MASK), invc, y0);
    // log2: 8 fp64 ops (the logc + k sum is folded into the table), then y*log2(x)
    const double z = (double)__uint_as_float(iz);
    const double r = __fma_rn(z, invc, -1.0);
    double y = __fma_rn(r, pk[0], pk[1]);
    const double p = __fma_rn(r, pk[2], pk[3]);
    const double r2 = __dmul_rn(r, r);
    double q = __fma_rn(r, pk[4], y0);
    const double r4 = __dmul_rn(r2, r2);
    q = __fma_rn(r2, p, q);
    y = __fma_rn(y, r4, q);
    const double xd = __dmul_rn(pk[5], y);
    // exp2: 8 fp64 ops
    double kd = __dadd_rn(xd, pk[9]);
    const uint32_t ki = (uint32_t)__double2loint(kd);
    kd = __dadd_rn(kd, pk[10]);
    const double rr = __dadd_rn(xd, -kd);
    uint32_t tlo, thi;
    lds_u32x2(ta.exp | ((ki << EXP_SHIFT) & EXP_MASK), tlo, thi);
    thi += ki << 15;  // t += ki << 47
    const double sc = __hiloint2double((int)thi, (int)tlo);
    const double zz = __fma_rn(rr, pk[6], pk[7]);
    const double rr2 = __dmul_rn(rr, rr);
    double yy = __fma_rn(rr, pk[8], 1.0);
    yy = __fma_rn(zz, rr2, yy);
    yy = __dmul_rn(yy, sc);
    return __double2float_rn(yy);
}

// ---- the look-back row pipeline ---------------------------------------------------------------------------------
// One look-back row = the LB candidates (dq = 1..LB) of one dj: LB independent powf(discrep, 1.2f) chains, replayed
// BREADTH-FIRST (every stage applied to all chains before the next one) so that dependent fp64 instructions of a
// chain are LB issue slots apart (DFMA latency ~10 cycles, one fp64 warp-instruction per 2 cycles per SM
// sub-partition).  The row is split in two so that the kernel can software-pipeline rows: row_prefetch() turns the
// discrepancies into table operands (address chain + LDS.128 + F2F, ~35 cycles of latency) one row AHEAD of
// row_powf(), which is pure fp64 work on operands that are already in registers.  Same arithmetic as powf12_fast.
struct PowRow {
    double z[LB], invc[LB], y0[LB];
};

// db_row = this lane's b[j] - b[j-dj] (or x-side value when the maps are swapped by the caller -- they are not:
// swap only changes the fn/fp terms); dx[c] = x[q] - x[q-1-c].  umax accumulates the out-of-domain flag.
__device__ __forceinline__ void row_prefetch(PowRow &n, float db_row, const float (&dx)[LB], const TabAddr ta,
                                             uint32_t &umax) {
    uint32_t u[LB];
#pragma unroll
    for (int c = 0; c < LB; c++) {
        const uint32_t ix = __float_as_uint(fabsf(__fsub_rn(db_row, dx[c])));
        u[c] = ix - IX_LO;
        const uint32_t iz = ix - (u[c] & 0xff800000u) + ((uint32_t)KLO << 23);
        lds_f64x2(ta.t2d | ((u[c] >> T2D_SHIFT) & T2D_MASK), n.invc[c], n.y0[c]);
        n.z[c] = f2d_normal(iz);
    }
    umax = __vimax3_u32(umax, __vimax3_u32(u[0], u[1], u[2]), __vimax3_u32(u[3], u[4], u[5]));
}

// Where the fp64 constants of the replay come from: SA_PK_LIT 1 = literals (ptxas picks immediates / uniform registers /
// its literal constant bank), 0 = the kernel-parameter array pk[] (constant bank 0).
#ifndef SA_PK_LIT
#define SA_PK_LIT 1
#endif
#if SA_PK_LIT
__device__ __forceinline__ constexpr double pk_lit(int i) {
    return i == 0 ? SA_POWF_A0 : i == 1 ? SA_POWF_A1 : i == 2 ? SA_POWF_A2 : i == 3 ? SA_POWF_A3 : i == 4 ? SA_POWF_A4
         : i == 5 ? SA_Y12 : i == 6 ? SA_EXP2F_C0 : i == 7 ? SA_EXP2F_C1 : i == 8 ? SA_EXP2F_C2
         : i == 9 ? SA_EXP2F_SHIFT : -SA_EXP2F_SHIFT;
}
#define PK(i) pk_lit(i)
#else
#define PK(i) pk[i]
#endif
__device__ __forceinline__ void row_powf(const PowRow &w, const TabAddr ta, const double (&pk)[12], float (&out)[LB]) {
    constexpr int NC = LB;
    uint32_t ki[NC], tlo[NC], thi[NC];
    double r[NC], y[NC], p[NC], r2[NC], q[NC], xd[NC], kd[NC], rr[NC], zz[NC], yy[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) r[c] = __fma_rn(w.z[c], w.invc[c], -1.0);
#pragma unroll
    for (int c = 0; c < NC; c++) y[c] = __fma_rn(r[c], PK(0), PK(1));
#pragma unroll
    for (int c = 0; c < NC; c++) p[c] = __fma_rn(r[c], PK(2), PK(3));
#pragma unroll
    for (int c = 0; c < NC; c++) r2[c] = __dmul_rn(r[c], r[c]);
#pragma unroll
    for (int c = 0; c < NC; c++) q[c] = __fma_rn(r[c], PK(4), w.y0[c]);
#pragma unroll
    for (int c = 0; c < NC; c++) q[c] = __fma_rn(r2[c], p[c], q[c]);
#pragma unroll
    for (int c = 0; c < NC; c++) r2[c] = __dmul_rn(r2[c], r2[c]);  // r4
#pragma unroll
    for (int c = 0; c < NC; c++) y[c] = __fma_rn(y[c], r2[c], q[c]);
#pragma unroll
    for (int c = 0; c < NC; c++) xd[c] = __dmul_rn(PK(5), y[c]);
#pragma unroll
    for (int c = 0; c < NC; c++) kd[c] = __dadd_rn(xd[c], PK(9));
#pragma unroll
    for (int c = 0; c < NC; c++) {
        ki[c] = (uint32_t)__double2loint(kd[c]);
        lds_u32x2(ta.exp | ((ki[c] << EXP_SHIFT) & EXP_MASK), tlo[c], thi[c]);
    }
#pragma unroll
    for (int c = 0; c < NC; c++) kd[c] = __dadd_rn(kd[c], PK(10));
#pragma unroll
    for (int c = 0; c < NC; c++) rr[c] = __dadd_rn(xd[c], -kd[c]);
#pragma unroll
    for (int c = 0; c < NC; c++) zz[c] = __fma_rn(rr[c], PK(6), PK(7));
#pragma unroll
    for (int c = 0; c < NC; c++) yy[c] = __fma_rn(rr[c], PK(8), 1.0);
#pragma unroll
    for (int c = 0; c < NC; c++) rr[c] = __dmul_rn(rr[c], rr[c]);
#pragma unroll
    for (int c = 0; c < NC; c++) yy[c] = __fma_rn(zz[c], rr[c], yy[c]);
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const double sc = __hiloint2double((int)(thi[c] + (ki[c] << 15)), (int)tlo[c]);
        out[c] = d2f_normal(__dmul_rn(yy[c], sc));
    }
}

// powf(ax, 1.2f) for any ax >= 0: the table replay inside the fast domain, the generic path outside it.
__device__ __noinline__ float powf12_any(float ax, uint32_t t2d, uint32_t exp) {
    TabAddr ta;
    ta.t2d = t2d, ta.exp = exp;
    uint32_t u;
    const float f = powf12_fast(ax, ta, c_pk, u);
    return (u < IX_RANGE) ? f : powf12_generic(ax, c_logtab, reinterpret_cast<const uint64_t *>(c_exptab));
}

// ---- exact pruning -------------------------------------------------------------------------------------------------
// Only the maximum over the 36 candidates of a cell is needed (SegAligner.cpp:123-127), and it is needed bit-exactly.
// Every fp32 operation of `S[i][p] + (10000 - ((fn + fp) + delta))` is monotone in delta, so a LOWER bound of
// delta gives an UPPER bound `hi` of the candidate's exact score with no error analysis.  Per cell:
//   1. all 36 candidates get `hi` from delta_lo(d) = ex2.approx(1.2 * lg2.approx(d) - PRUNE_C) (two MUFU, fp32 only;
//      delta_lo <= the bit-exact powf replay for EVERY float, certified exhaustively on the GPU by
//      sa_check_prune_bound); the largest `hi` (with its candidate number) and the second largest are tracked;
//   2. the candidate with the largest `hi` is evaluated exactly (the one fp64 replay of the cell);
//   3. if its exact score is larger than the second largest `hi` it is the strict maximum and the cell is done;
//      otherwise (ties and near-ties: rare) the whole cell is redone exactly by cell_slow().
// Measured on the C2 workload: step 3 settles 99.98 % of the cells.
#ifndef SA_PRUNE_C
#define SA_PRUNE_C 0x1p-16f  // >= 1.316e-5 needed for every float (measured, tools/chk_prune.py); 2^-16 = 1.526e-5
#endif
constexpr float PRUNE_C = SA_PRUNE_C;  // log2-domain safety margin of delta_lo (MUFU errors are ~2^-22)
__device__ __forceinline__ float delta_lo_approx(float d) {
    float a, e;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(fabsf(d)));
    const float t = __fmaf_rn(a, 1.2f, -PRUNE_C);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return e;
}

// Packed fp32 (sm_100 add/sub/fma .f32x2: two IEEE round-to-nearest operations per instruction, bit-identical to the scalar
// ones).  The bound pass evaluates candidates in pairs; packing its five chain operations halves their issue slots, the
// kernel's second ceiling next to the MUFU pipe.  SA_PACKED 0 keeps the scalar chain (A/B timing).
#ifndef SA_PACKED
#define SA_PACKED 1
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// powf(ax, 1.2f), inline fast path with the generic path behind a (rare) branch
__device__ __forceinline__ float powf12_exact(float ax, const TabAddr ta, const double (&pk)[12]) {
    uint32_t u;
    float f = powf12_fast(ax, ta, pk, u);
    if (u >= IX_RANGE) f = powf12_slow(ax);
    return f;
}

// Whole-cell exact evaluation in the reference's candidate order and tie rule (scan dj, dq ascending, replace on >=:
// the maximum, ties -> the LAST candidate in scan order).  Used by the dense kernel for cells with an out-of-domain
// discrepancy and by the pruning kernel for (near-)ties; starts from an already exact (best, code) when the caller
// has one and then skips the fp64 replay of every candidate whose upper bound cannot reach the running maximum.
// a_row = shared address of this lane's row operands (db[0..5], fb[0..5], 128 B apart); f[] = fx (non-swap: fp_t by
// column look-back; with swap the row-side fb is read from a_row).
struct CellRes {
    float best;
    int code;
};
__device__ __noinline__ CellRes cell_slow(bool swap, uint32_t a_row, float dx0, float dx1, float dx2, float dx3,
                                          float dx4, float dx5, float f0, float f1, float f2, float f3, float f4,
                                          float f5, uint32_t pbase, uint32_t t2d, uint32_t exp, float best0, int code0,
                                          float gain = 10000.0f) {
    const float dx[LB] = {dx0, dx1, dx2, dx3, dx4, dx5};
    const float f[LB] = {f0, f1, f2, f3, f4, f5};
    CellRes res;
    res.best = best0;
    res.code = code0;
    for (int dj = 1; dj <= LB; dj++) {
        const float dbv = lds_f32(a_row + (dj - 1) * 128);
        const float fbv = swap ? lds_f32(a_row + (LB + dj - 1) * 128) : 0.0f;
        for (int dq = 1; dq <= LB; dq++) {
            const float sc = lds_f32(pbase + (LB - dq) * (40 * 4) - dj * 4);  // mirrored ring, see PC() in fill_kernel
            const float d = __fsub_rn(dbv, dx[dq - 1]);
            const float fn = 5000.0f * (float)((swap ? dq : dj) - 1);
            const float fp = swap ? fbv : f[dq - 1];
            const float fnfp = __fadd_rn(fn, fp);
#if SA_PRUNE
            const float hi = __fadd_rn(sc, __fsub_rn(gain, __fadd_rn(fnfp, delta_lo_approx(d))));
            if (!(hi >= res.best)) continue;  // its exact score is <= hi < the running maximum
#endif
            const float delta = powf12_any(fabsf(d), t2d, exp);
            const float ns = __fadd_rn(sc, __fsub_rn(gain, __fadd_rn(fnfp, delta)));
            const int ck = 1 + (dj - 1) * LB + (dq - 1);
            if (ns > res.best || (ns == res.best && ck >= res.code)) {
                res.best = ns;
                res.code = ck;
            }
        }
    }
    return res;
}

// ---- the per-pair alignment loop on the device ("rounds", STORE_MODE 3) ---------------------------------------------------
// After a (full or windowed) sweep of the pair's matrix: sg_aln_backtrack_start (SegAligner.cpp:172-195) over the persisted
// last row and last two columns, get_aln_list (SAHelper.cpp:83-95) over the persisted predecessor codes with the scores
// along the path recomputed exactly (as trace_kernel does), compute_partial_score_threshold (SegAligner.cpp:23-36), the
// loop bookkeeping of run_SA_aln (main.cpp:268-300), and the window of strips the NEXT round has to re-sweep.
// Why a window is exact: dp_align's cell (j, q) reads rows j-6..j-1 of columns < q only, so by induction it depends on rows
// >= j - 6q.  Masking the contig labels of a path that spans rows [j_lo, j_hi] therefore cannot change any row < j_lo nor
// any row > j_hi + 6 (nx - 1); rows in between are re-swept strip by strip, the strip above the window supplying its
// (unchanged) bottom rows from the per-strip halo slots.
// Called by all 32 lanes of the warp; returns true when the pair is finished.  count / sb / se are warp-uniform.
__device__ __noinline__ bool rounds_finish(const FillParams &P, long long gunit, long long pidx, int contig, int nb, int seg, int nx,
                                           float thr, int round, int &count, int &sb, int &se) {
    const int lane = threadIdx.x & 31;
    const float *col2 = P.r_col2 + gunit * 2 * P.r_max_nb;
    const float *lastrow = P.r_lastrow + gunit * P.r_max_nx;
    // backtrack start: scan position e = the reference's scan order (last row by q, then rows by j over the last 2 columns);
    // strict '>' keeps the first maximum; -inf cells are never selected
    float st_s = NEG_INF_F;
    int st_e = -1;
    const int n_e = nx + 2 * nb;
    for (int e = lane; e < n_e; e += 32) {
        const float v = (e < nx) ? __ldcg(lastrow + e) : __ldcg(col2 + (e - nx));
        if (v > st_s) {
            st_s = v;
            st_e = e;
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, st_s, m);
        const int oe = __shfl_xor_sync(0xffffffffu, st_e, m);
        if (oe >= 0 && (st_e < 0 || os > st_s || (os == st_s && oe < st_e))) {
            st_s = os;
            st_e = oe;
        }
    }
    int st_j = -1, st_q = -1;
    if (st_e >= 0) {
        if (st_e < nx) st_j = nb - 1, st_q = st_e;
        else st_j = (st_e - nx) >> 1, st_q = nx - 2 + ((st_e - nx) & 1);
    }
    int stop = 1, n_sb = 0, n_se = 0, n_count = count;
    if (lane == 0) {
        int n = 0;
        long long aoff = 0;
        float exp_thresh = 0.0f;
        int32_t *pj = P.r_path + gunit * 2 * P.r_max_path, *pq = pj + P.r_max_path;
        float *ps = P.r_ps + gunit * P.r_max_path;
        if (st_e >= 0) {
            const uint8_t *C = P.r_code + gunit * P.r_cells;
            int j = st_j, q = st_q;
            for (;;) {  // get_aln_list: follow `previous` until {-1,-1}
                pj[n] = j;
                pq[n] = q;
                n++;
                const int cd = __ldcg(C + (long long)j * nx + q);
                if (cd == 0 || n >= P.r_max_path) break;
                j -= (cd - 1) / LB + 1;
                q -= (cd - 1) % LB + 1;
            }
            const int64_t boff = P.contigs.pos_off[contig], xoff = P.segs.pos_off[seg];
            // compute_partial_score_threshold (SegAligner.cpp:23-36) on the path; aln_list is end -> start
            const float *cpos = P.contigs.pos + boff;  // nb + 1 entries, the last one is the contig length
            const int first_lab = pj[n - 1], last_lab = pj[0];
            const float by_lab = __fmul_rn(thr, __fdiv_rn((float)(last_lab - first_lab + 1), (float)nb));
            const float by_len = __fmul_rn(thr, __fdiv_rn(__fsub_rn(cpos[last_lab], cpos[first_lab]), cpos[nb]));
            exp_thresh = fmaxf(by_lab, by_len);  // main.cpp:290
            // Does the traceback leave the device?  Always without a filter; with one only when the alignment was accepted and
            // passes the file filters that need just its ends, length and total score (main.cpp:303-325; the mean with slack,
            // the caller redoes it exactly): everything else the caller's replay reads is in the round record.
            bool keep = true;
            if (P.r_flt_on) {
                keep = st_s > exp_thresh && n >= P.r_flt_min_len;
                if (keep && P.r_flt_tip && (nb - 1) - last_lab > 3 && first_lab > 2) keep = false;  // touches neither contig end
                if (keep) {
                    const float mean = __fdiv_rn(st_s, (float)(n - 1));
                    keep = mean >= 0.999f * (n < 5 ? P.r_flt_mean_short : P.r_flt_mean);
                }
            }
            aoff = -1;
            if (keep) {
                // scores along the path, start -> end, exactly as dp_align produced them (see trace_kernel)
                const LabelOps *bops = P.contigs.ops + (boff - contig);
                const LabelOps *xops = P.segs.ops + (xoff - seg);
                float sv = (pj[n - 1] == 0 || pq[n - 1] < 2) ? 0.0f : NEG_INF_F;  // init_semiglobal_aln
                ps[n - 1] = sv;
                int jp = pj[n - 1], qp = pq[n - 1];
                for (int t = n - 2; t >= 0; t--) {
                    const int jc = pj[t], qc = pq[t];
                    const int dj = jc - jp, dq = qc - qp;
                    const float discrep = fabsf(__fsub_rn(bops[jc].d[dj - 1], xops[qc].d[dq - 1]));
                    const float delta = powf12_slow(discrep);
                    const float fnfp = (dj == 1) ? xops[qc].f[dq - 1] : __fadd_rn(5000.0f * (float)(dj - 1), xops[qc].f[dq - 1]);
                    sv = __fadd_rn(sv, __fsub_rn(10000.0f, __fadd_rn(fnfp, delta)));
                    ps[t] = sv;
                    jp = jc;
                    qp = qc;
                }
                aoff = (long long)atomicAdd(P.r_arena_top, (unsigned long long)n);
                if (aoff + n <= P.r_arena_cap) {
                    for (int t = 0; t < n; t++) {
                        P.r_pj[aoff + t] = pj[t];
                        P.r_pq[aoff + t] = pq[t];
                        P.r_psc[aoff + t] = ps[t];
                    }
                } else {
                    *P.r_overflow = 1;
                }
            }
        }
        sa_round rec;
        rec.score = st_s;
        rec.j = st_j;
        rec.q = st_q;
        rec.path_len = n;
        rec.path_off = aoff;
        rec.j_first = n > 0 ? pj[n - 1] : -1;
        rec.reserved = 0;
        P.r_rounds[pidx * P.r_max_alns + round] = rec;
        if (n > 0 && st_s > exp_thresh) {  // main.cpp:295: accepted -> count, the path's contig labels join the used set
            n_count = count + 1;
            uint32_t *ub = P.r_used + gunit * P.r_used_words;
            for (int t = 0; t < n; t++) {
                const int jj = pj[t];
                __stcg(ub + (jj >> 5), __ldcg(ub + (jj >> 5)) | (1u << (jj & 31)));
            }
            // main.cpp:268: another round, over the rows this path can influence -- unless the caller has declared rounds that
            // score below r_flt_stop_below unobservable (later rounds cannot score higher than this one)
            if (n_count < P.r_max_alns && !(st_s < P.r_flt_stop_below)) {
                stop = 0;
                const int j_lo = pj[n - 1], j_hi = pj[0];
                const int nstrips = (nb + 31) >> 5;
                n_sb = j_lo >> 5;
                n_se = ((j_hi + LB * (nx - 1)) >> 5) + 1;
                if (n_se > nstrips) n_se = nstrips;
            }
        }
        if (stop) P.r_n_rounds[pidx] = round + 1;
    }
    __syncwarp();
    stop = __shfl_sync(0xffffffffu, stop, 0);
    count = __shfl_sync(0xffffffffu, n_count, 0);
    sb = __shfl_sync(0xffffffffu, n_sb, 0);
    se = __shfl_sync(0xffffffffu, n_se, 0);
    return stop != 0;
}

// Tip pass only (sa_rounds_filter.tip_windows): can ANY round of this pair write a tip alignment?  run_SA_aln writes one only if
// its first label is <= 2 or its last label >= nb - 4 (main.cpp:303-308), it has >= min_len labels and its mean score per step
// reaches a bar (main.cpp:310-325).  Three facts make a cheap, exact pre-check out of that:
//  * A path takes one segment label per step and 1 + (skipped contig labels) rows; a skipped label costs 5000 (fn_t), and a path
//    that passes the mean filter can afford c = 10000 - bar per step on average, so it spans at most (nx - 1)(1 + c / 5000)
//    rows (win_span: 1.25 nx for the tip pass instead of the 6 nx a path could span at all).  One that starts in rows 0..2
//    ends in rows <= 2 + span; one that ends in rows >= nb - 4 lies in rows >= nb - 4 - span.  The end of a round's path is
//    its backtrack start: a cell of the last row or of the last two columns.  A window swept on its own (the rows above it
//    taken as absent) gives every cell the best path into it that lies INSIDE the window, which is all the bound needs.
//  * A cell's value is the best path into it, `val = max(init, best candidate)`, every operation monotone: masking more labels
//    (later rounds) can only lower it, and so can closing start cells.  A path that starts in rows 0..2 starts in row 0 or in
//    columns 0, 1 of rows 1, 2 (the cells init_semiglobal_aln sets to 0); the top strips are swept with the free starts of
//    columns 0, 1 below row 2 closed, so that what reaches the last two columns there is the best path FROM THE CONTIG END --
//    otherwise every ordinary alignment that merely ends inside the window keeps the pair alive (measured on C2: 26 % of the
//    pairs survive a plain score bound over 6 nx rows, 16 % with the starts closed, 4.6 % with the gain below as well, 0.9 %
//    with the windows cut to win_span; profiles/r2_c2_whole_run_tip_windows*.txt).
//  * Both windows are swept with a per-step gain G < 10000 (r_win_gain, set by the host from the filter's mean bar): a path that
//    passes the mean filter has penalties <= 10000 - bar per step on average and is worth >= (G - 10000 + bar)(n - 1) > 0 here,
//    while chance alignments, whose penalties exceed G on most steps, sink below zero.  (A plain score bound cannot tell a
//    40-label noise alignment of 45 000 from a 6-label tip alignment of 50 000.)
// The maximum over the windows' backtrack-start cells is compared with r_win_min (a fraction of what the shortest writable path
// is worth); below it, the pair is finished without ever having been aligned.  All 32 lanes call this.
// contig labels a writable path can span beyond its first: one per step plus the labels it skips, and skipping costs 5000 each
// (fn_t, SegAligner.cpp:89) out of a penalty budget of c per step, so at most (nx - 1)(1 + c / 5000) (r_win_slope); never more
// than 6 per step
__device__ __forceinline__ int win_span(const FillParams &P, int nx) {
    return min((int)ceilf((float)(nx - 1) * P.r_win_slope) + 1, LB * (nx - 1));
}
__device__ __noinline__ float end_window_best(const FillParams &P, long long gunit, int nb, int nx) {
    const int lane = threadIdx.x & 31;
    const float *col2 = P.r_col2 + gunit * 2 * P.r_max_nb;
    const float *lastrow = P.r_lastrow + gunit * P.r_max_nx;
    float m = NEG_INF_F;
    const int top = min(nb, 3 + win_span(P, nx));
    for (int e = lane; e < 2 * top; e += 32) m = fmaxf(m, __ldcg(col2 + e));
    const int bot = max(0, nb - 4);
    for (int e = 2 * bot + lane; e < 2 * nb; e += 32) m = fmaxf(m, __ldcg(col2 + e));
    for (int e = lane; e < nx; e += 32) m = fmaxf(m, __ldcg(lastrow + e));
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
    return m;
}

constexpr int RING_W = 40;  // 6 halo rows + 32 lanes + 2 pad
constexpr int CPL_LAG = 2;  // coupled units: columns a consumer stays behind its producer
// STORE staging per warp: scores [32][9] floats + codes [32][12] bytes + the problem's cell offset (8 B, read back at every
// flush from shared memory instead of living in registers or being re-fetched from global memory)
constexpr int STAGE_BYTES_PER_WARP = 32 * 9 * 4 + 32 * 12 + 16;
#define NEG_INF (-CUDART_INF_F)

// NW = warps that cooperate on one problem.
//   NW == 1 : every warp of the CTA pulls its own problem (bulk of the work: thousands of small problems).
//   NW  > 1 : the whole CTA (NW warps) sweeps ONE large problem as a wavefront: warp w owns row strips
//             w, w+NW, ... and runs one column behind warp w-1 (strip s at column q needs the bottom rows of
//             strip s-1 at columns < q).  One __syncthreads per step; bottom rows travel through a
//             double-buffered shared exchange (warp w-1 -> w) and, for the wrap-around NW-1 -> 0, through the
//             L2-resident halo scratch.  A round takes R = max(nx, NW) steps so the wrap-around is never early.
//   CS  > 1 : a thread-block CLUSTER of CS such CTAs (CS*NW warps on CS SMs) sweeps one huge problem: same wavefront,
//             the step barrier becomes a cluster barrier, bottom rows cross CTA boundaries through a small
//             L2-resident mailbox.  Used for chromosome-sized problems whose single-SM time would be the tail.
// STORE_MODE: 0 = nothing per cell (scoring), 1 = predecessor codes only (alignment), 2 = codes + scores + per-row maxima
// (detection, -nl-free test hook).  Separate instantiations keep the alignment kernel free of the registers mode 2 needs.
// STORE_MODE 3 = "rounds" (NW == 1 only): the warp keeps its pair and runs the whole run_SA_aln loop on it (main.cpp:262-300):
// codes go to per-worker scratch, every strip's bottom rows / the last row / the last two columns persist in HBM, and each
// later round re-sweeps only the strips its predecessor's path can influence (see rounds_finish below).
// CPL ("coupled units", NW == 1): the way large problems are swept since round 2.  A problem is cut into U units; unit u owns
//             the row strips u, u+U, ... and is processed by whichever warp pulls it from the queue (units of a problem are
//             queued consecutively, largest problems first).  Strip s needs the bottom 6 rows of strip s-1 column by column:
//             the producing warp writes them as 8-byte {value, tag} entries (tag = s * nx + q + 1, unique within a launch)
//             into its unit's mailbox in L2, the consuming warp issues the load at the start of the step and only looks at
//             it at the end, re-polling while the tag is not the expected one.  No barrier, no fence (an entry is one
//             8-byte store, self-validating), no co-scheduling constraint beyond U <= resident warps: a straggler delays
//             only its successor, and every warp runs the bulk code at the bulk rate.  Overwriting is safe because unit u
//             reaches strip s+U only after the whole ring of units, including its consumer, has passed strip s.
template <int STORE_MODE, bool SWAP, int NW, int MINB, int CS, bool CPL = false>
__global__ void __launch_bounds__((NW == 1 ? FILL_WARPS : NW) * 32, MINB) fill_kernel(const FillParams P) {
    constexpr bool STORE = STORE_MODE != 0;
    constexpr bool STORE_S = STORE_MODE == 2;
    constexpr bool ROUNDS = STORE_MODE == 3;
    static_assert(!ROUNDS || (NW == 1 && CS == 1 && !SWAP), "the rounds variant is one warp per pair, semi-global, no swap");
    static_assert(!CPL || (NW == 1 && CS == 1 && !ROUNDS), "coupled units are single warps");
    constexpr int WPC = (NW == 1) ? FILL_WARPS : NW;  // warps per CTA
    constexpr int NWT = NW * CS;                        // warps in the wavefront
    static_assert(CS == 1 || NW > 1, "clusters only for the cooperative shape");
    const int crank = (CS > 1) ? (int)cg::this_cluster().block_rank() : 0;
    extern __shared__ unsigned char s_dyn[];
    PowTabSmem &s_tab = *tab_place(s_dyn);
    // dynamic shared memory behind the table region: the mirrored ring (column q lives in slots q%6 and q%6+6, see
    // PC()) and the lane-private row operands db[0..5], fb[0..5] (LabelOps of the lane's row)
    float(*s_ring)[2 * LB][RING_W] = reinterpret_cast<float(*)[2 * LB][RING_W]>(s_dyn + TAB_SMEM_BYTES);
    float(*s_rowop)[2 * LB][32] =
        reinterpret_cast<float(*)[2 * LB][32]>(s_dyn + TAB_SMEM_BYTES + WPC * 2 * LB * RING_W * sizeof(float));
    // STORE only: per-warp staging tiles of 32 rows x 8 columns (scores: row pitch 9 floats, codes: row pitch 12 bytes) so
    // that the stored matrices are written 8 consecutive columns (32 B / 8 B) per row instead of one element per row
    unsigned char *s_stage = s_dyn + TAB_SMEM_BYTES + WPC * (2 * LB * RING_W + 2 * LB * 32) * sizeof(float);
    __shared__ float s_xbuf[WPC][2][8];
    __shared__ unsigned long long s_next;
    __shared__ float s_red_s[WPC];
    __shared__ int s_red_j[WPC], s_red_q[WPC];
    __shared__ long long s_red_o[WPC];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    powtab_init(s_tab);
    __syncthreads();
    const TabAddr a_tab = tab_addr(s_tab);
    // byte address of this lane's own entry in ring slot 0
    const uint32_t a_my = (uint32_t)__cvta_generic_to_shared(&s_ring[warp][0][6 + lane]);
    const uint32_t a_halo = (uint32_t)__cvta_generic_to_shared(&s_ring[warp][0][lane]);  // lanes 0..5 only
    const uint32_t a_row = (uint32_t)__cvta_generic_to_shared(&s_rowop[warp][0][lane]);
    const uint32_t a_tile_s = (uint32_t)__cvta_generic_to_shared(s_stage + warp * STAGE_BYTES_PER_WARP);  // [32][9] floats
    const uint32_t a_tile_c = a_tile_s + 32 * 9 * 4;                                                    // [32][12] bytes

    const long long gunit = (NW == 1) ? ((long long)blockIdx.x * FILL_WARPS + warp) : (long long)(blockIdx.x / CS);
    // rounds: one slot per strip ([strip][q][8]) in the worker's scratch, so a later round finds the rows above its window
    float *halo = ROUNDS ? P.r_halo + gunit * P.r_halo_floats : P.halo + gunit * P.halo_stride;
    const int w = (NW == 1) ? 0 : crank * NW + warp;  // position in the wavefront
    // cluster mailbox (CS > 1): [0..CS) next-problem broadcast slot 0, then per-CTA exchange [CS][2][8] floats,
    // then per-CTA start candidates
    float *cmail = (CS > 1) ? P.cluster_mail + gunit * CLUSTER_MAIL_FLOATS : nullptr;

    for (;;) {
        unsigned long long t_idx = 0;
        if (NW == 1) {
            if (lane == 0) t_idx = atomicAdd(P.counter, 1ull);
            t_idx = __shfl_sync(0xffffffffu, t_idx, 0);
        } else if (CS == 1) {
            __syncthreads();
            if (threadIdx.x == 0) s_next = atomicAdd(P.counter, 1ull);
            __syncthreads();
            t_idx = s_next;
        } else {
            cg::this_cluster().sync();
            if (crank == 0 && threadIdx.x == 0) {
                const unsigned long long v = atomicAdd(P.counter, 1ull);
                __stcg(reinterpret_cast<unsigned long long *>(cmail), v);
            }
            cg::this_cluster().sync();
            t_idx = __ldcg(reinterpret_cast<const unsigned long long *>(cmail));
        }
        if (t_idx >= (unsigned long long)P.n_problems) break;
        int cu = 0, cU = 1, cslot = 0;  // coupled: this warp's unit, the units of the problem, its first mailbox slot
        long long pidx;
        if (CPL) {
            const CplUnit un = P.units[t_idx];
            pidx = un.problem;
            cu = un.u;
            cU = un.U;
            cslot = un.slot0;
        } else {
            pidx = P.order ? P.order[t_idx] : (long long)t_idx;
        }
#ifdef SA_CPL_DEBUG
        if (CPL && lane == 0) P.cpl_done[P.n_problems + 64 + cslot + cu] = 0x40000000;  // claimed
#endif
        const unsigned long long *mb_in = CPL ? P.mbox + (long long)(cslot + (cu + cU - 1) % cU) * P.mbox_stride : nullptr;
        unsigned long long *mb_out = CPL ? P.mbox + (long long)(cslot + cu) * P.mbox_stride : nullptr;
        const sa_problem pr = P.problems[pidx];
        const int64_t boff = P.contigs.pos_off[pr.contig];
        const int nb = (int)(P.contigs.pos_off[pr.contig + 1] - boff) - 1;
        // label / bitmap bases as 32-bit element offsets from the kernel-parameter pointers (one register each instead
        // of a 64-bit pointer; sa_set_maps rejects map sets with 2^31 or more entries)
        const int bo = (int)(boff - pr.contig);  // bops = P.contigs.ops + bo: only needed at strip set-up
        const int64_t xoff = P.segs.pos_off[pr.seg];
        const int nx = (int)(P.segs.pos_off[pr.seg + 1] - xoff) - 1;
        const LabelOps *xops = P.segs.ops + (xoff - pr.seg);
        const int used_off = (P.used_bits && pr.used_slot >= 0) ? pr.used_slot * (int)P.used_words_per_slot : -1;  // host-checked < 2^31
        // (STORE: the output pointers are rebuilt at every flush from P.S / P.code + the cell offset parked in shared memory,
        // instead of living in four registers through the bound pass -- the kernel is compiled for 80 registers)
        if (STORE) {
            if (lane == 0) {
                const long long coff = ROUNDS ? gunit * P.r_cells : P.cell_off[pidx];
                asm volatile("st.shared.s64 [%0], %1;" ::"r"(a_tile_c + 32 * 12), "l"(coff) : "memory");
            }
            __syncwarp();
        }

        unsigned n_replay = 0;  // pruning statistics of this lane: fp64 replays (rescans are rare: counted by an atomic)
        // backtrack-start tracking (per lane, reduced at the end)
        float st_s = NEG_INF;
        int st_j = -1, st_q = -1;
        // position of a cell in sg_aln_backtrack_start's scan (last row by q, then rows by j over the last 2 columns);
        // derived from (j, q) when needed instead of being carried in two registers through the sweep
        auto sg_ord = [&](int jj, int qq) -> long long {
            if (jj < 0) return 0x7fffffffffffffffll;
            return (jj == nb - 1) ? (long long)qq : (long long)nx + 2ll * jj + (qq - (nx - 2));
        };
        long long st_ord = 0x7fffffffffffffffll;  // filled in after the sweep

        const int nstrips = (nb + 31) >> 5;
        const int R = (NW == 1) ? nx : (nx > NWT ? nx : NWT);  // steps per round
        const int rounds = (nstrips + NWT - 1) / NWT;
        // rounds variant: strips [r_sb, r_se) are swept in this pass; accepted alignments so far; pass number
        int r_sb = 0, r_se = nstrips, r_count = 0, r_round = 0;
        // tip windows: 1 = sweeping the top strips, 2 = sweeping the bottom strips (started cold), 0 = the loop proper
        int r_pre = 0;
        if (ROUNDS && P.r_flt_tipwin) {
            const int span = win_span(P, nx);
            const int s_top = (3 + span + 31) >> 5;  // strips that hold rows 0 .. 2 + span
            const int bot_row = nb - 4 - span;       // first row of a writable path that ends in rows >= nb - 4
            if (bot_row > 0 && (bot_row >> 5) >= s_top) {      // the two windows are disjoint: worth a look before the full sweep
                r_pre = 1;
                r_se = s_top;
            }
        }
        if (ROUNDS) {
            uint32_t *ub = P.r_used + gunit * P.r_used_words;  // the pair's used labels: a copy of its seed bitmap (or none)
            for (int i = lane; i < ((nb + 31) >> 5); i += 32) __stcg(ub + i, used_off >= 0 ? P.used_bits[used_off + i] : 0u);
            __syncwarp();
            const float thr0 = P.r_thr[pidx];
            if (!(thr0 >= thr0)) {  // NaN threshold: `curr_best_score >= exp_thresh` is false from the start (main.cpp:268)
                if (lane == 0) P.r_n_rounds[pidx] = 0;
                continue;
            }
        }
        for (;;) {  // one pass, except in the rounds variant
        // < 2^31: check_problems() rejects problems with more than 2^35 cells
        // what a step is worth before its penalties (score_f's 10000, SegAligner.cpp:92); the end windows of the tip pass are
        // swept with a smaller value: see end_window_best
        const float gain = (ROUNDS && r_pre != 0) ? P.r_win_gain : 10000.0f;
        const int T = ROUNDS ? (r_se - r_sb) * nx : CPL ? ((nstrips - cu + cU - 1) / cU) * nx : rounds * R + (NWT - 1);

        // per-warp sweep state
        int strip = ROUNDS ? r_sb : CPL ? cu : w;
        int q = 0;  // column within the current round (>= nx: idle tail of a short round)
        int j = 0, jc = 0;
        bool row_ok = false, row_used = false;
#if SA_PRUNE
        float db[LB], fb[LB];  // row operands also in registers (pass 1 is fully unrolled; pass 2 reads the shared copy)
#pragma unroll
        for (int d = 0; d < LB; d++) db[d] = 0.0f, fb[d] = 0.0f;
#endif
        // Mirrored ring: the value of column q is written to slots s = q % 6 AND s + 6, so the six previous columns
        // q-1-c (c = 0..5) sit at the STATIC offsets (5 - c) from slot s: no pointer rotation, immediate LDS offsets.
        uint32_t pbase = a_my;  // a_my + slot * RING_W * 4
#define PC(c) (pbase + (uint32_t)((LB - 1 - (c)) * RING_W * 4))
        int slot = 0;
        float rm_s = 0.0f;  // detection: best positive score of this lane's row, ties -> larger q
        int rm_q = -1;

        for (int t = 0; t < T; t++) {
            if (ROUNDS && q == 0) {
                // Strips whose 32 rows are ALL used labels are not swept: dp_align leaves a used row at its initial values with
                // no predecessor (SegAligner.cpp:104-106), so the strip's outputs are known -- codes 0, and init_semiglobal_aln's
                // values (row 0 and columns 0, 1: 0, else -inf) in the strip-boundary rows, the last two columns and the last row.
                // A strip is one word of the pair's bitmap.  In the tip pass, where every contig label that an alignment of the
                // standard pass took is used from the start (main.cpp:597-602), this is almost half of all strips on C2.
                const uint32_t *ub = P.r_used + gunit * P.r_used_words;
                bool done = false;
                const bool open01 = r_pre != 1;  // columns 0, 1 are free starts in every row (not in the top end window)
                for (;;) {
                    if (strip >= r_se) {
                        done = true;
                        break;
                    }
                    const int rows_here = nb - (strip << 5) < 32 ? nb - (strip << 5) : 32;
                    const uint32_t valid = rows_here >= 32 ? 0xffffffffu : ((1u << rows_here) - 1u);
                    if ((__ldcg(ub + strip) & valid) != valid) break;
                    uint8_t *C = P.r_code + gunit * P.r_cells + (long long)(strip << 5) * nx;
                    for (int i = lane; i < rows_here * nx; i += 32) C[i] = 0;
                    if (rows_here == 32) {  // bottom rows for the strip below (rows 26..31 of this strip: never row 0)
                        float *h = halo + (long long)strip * nx * 8;
                        for (int i = lane; i < nx * 8; i += 32)
                            if ((i & 7) < 6) __stcg(h + i, ((i >> 3) < 2 && open01) ? 0.0f : NEG_INF);
                    }
                    const int jj = (strip << 5) + lane;
                    if (jj < nb) {
                        float *c2 = P.r_col2 + gunit * 2 * P.r_max_nb + 2 * jj;
                        const bool o01 = open01 || jj <= 2;
                        __stcg(c2, (jj == 0 || (nx - 2 < 2 && o01)) ? 0.0f : NEG_INF);
                        __stcg(c2 + 1, (jj == 0 || (nx - 1 < 2 && o01)) ? 0.0f : NEG_INF);
                    }
                    if (nb - 1 >= (strip << 5) && nb - 1 < (strip << 5) + 32) {
                        float *lr = P.r_lastrow + gunit * P.r_max_nx;
                        for (int i = lane; i < nx; i += 32) __stcg(lr + i, (nb - 1 == 0 || (i < 2 && (open01 || nb - 1 <= 2))) ? 0.0f : NEG_INF);
                    }
                    strip++;
                    t += nx;
                }
                __syncwarp();
                if (done) break;
            }
            const bool active = (NW == 1) ? true : ((t >= w) && (strip < nstrips) && (q < nx));  // NW==1: T = nstrips*nx
            float val = 0.0f, rv = 0.0f, hv = NEG_INF;
            unsigned long long hv_cpl = 0;  // coupled: the mailbox entry as loaded at the start of the step
            int code = 0;
            if (active) {
                if (q == 0) {  // strip set-up: row operands into registers, ring to -inf
                    if (CPL && strip > 0) {
                        // Start CPL_LAG columns behind the producer of the rows above: in steady state the entry for column q
                        // is then already in L2 when the step begins and the poll at its end never spins (measured before:
                        // 10 % of the warp time of the detection fill went into that spin, units running in lock-step).
                        const int ql = nx - 1 < CPL_LAG ? nx - 1 : CPL_LAG;
                        const uint32_t want = P.cpl_epoch | ((uint32_t)(strip - 1) * (uint32_t)nx + (uint32_t)ql + 1u);
                        if (lane == 0) {
                            unsigned long long e = 0;
                            for (unsigned spin = 0; spin < P.cpl_spin_limit; spin++) {
                                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(e) : "l"(mb_in + (long long)ql * 6) : "memory");
                                if ((uint32_t)(e >> 32) == want) break;
                                if ((spin & 1023u) == 1023u && *reinterpret_cast<volatile int32_t *>(P.cpl_error) != 0) break;
                            }
                        }
                        __syncwarp();
                    }
                    j = (strip << 5) + lane;
                    row_ok = j < nb;
                    jc = row_ok ? j : nb - 1;
                    const float4 *src = reinterpret_cast<const float4 *>(P.contigs.ops + bo + jc);
                    const float4 a = __ldg(src), b4 = __ldg(src + 1), c4 = __ldg(src + 2);
#if SA_PRUNE
                    db[0] = a.x, db[1] = a.y, db[2] = a.z, db[3] = a.w, db[4] = b4.x, db[5] = b4.y;
                    fb[0] = b4.z, fb[1] = b4.w, fb[2] = c4.x, fb[3] = c4.y, fb[4] = c4.z, fb[5] = c4.w;
#endif
                    // row operands go to lane-private shared memory: the look-back rows are a rolled loop (dynamic dj)
                    sts_f32(a_row + 0 * 128, a.x), sts_f32(a_row + 1 * 128, a.y), sts_f32(a_row + 2 * 128, a.z);
                    sts_f32(a_row + 3 * 128, a.w), sts_f32(a_row + 4 * 128, b4.x), sts_f32(a_row + 5 * 128, b4.y);
                    if (SWAP) {
                        sts_f32(a_row + 6 * 128, b4.z), sts_f32(a_row + 7 * 128, b4.w), sts_f32(a_row + 8 * 128, c4.x);
                        sts_f32(a_row + 9 * 128, c4.y), sts_f32(a_row + 10 * 128, c4.z), sts_f32(a_row + 11 * 128, c4.w);
                    }
                    if (ROUNDS) row_used = (__ldcg(P.r_used + gunit * P.r_used_words + (jc >> 5)) >> (jc & 31)) & 1u;
                    else row_used = used_off >= 0 ? ((P.used_bits[used_off + (jc >> 5)] >> (jc & 31)) & 1u) : false;
                    __syncwarp();
#pragma unroll
                    for (int s = 0; s < 2 * LB; s++) {
                        sts_f32(a_my + s * RING_W * 4, NEG_INF);
                        if (lane < 6) sts_f32(a_halo + s * RING_W * 4, NEG_INF);
                    }
                    __syncwarp();
                    slot = 0;
                    pbase = a_my;
                    rm_s = 0.0f;
                    rm_q = -1;
                }
                float dx[LB], fx[LB];
                {
                    const float4 *src = reinterpret_cast<const float4 *>(xops + q);
                    const float4 a = __ldg(src), b4 = __ldg(src + 1);
                    dx[0] = a.x, dx[1] = a.y, dx[2] = a.z, dx[3] = a.w, dx[4] = b4.x, dx[5] = b4.y;
                    if (!SWAP) {
                        const float4 c4 = __ldg(src + 2);
                        fx[0] = b4.z, fx[1] = b4.w, fx[2] = c4.x, fx[3] = c4.y, fx[4] = c4.z, fx[5] = c4.w;
                    }
                }
                // rows above the strip at column q (bottom rows of strip-1)
                if (strip > 0 && lane < 6) {
                    if (CPL) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(hv_cpl) : "l"(mb_in + (long long)q * 6 + lane) : "memory");
                    else if (ROUNDS) hv = (r_pre == 2 && strip == r_sb) ? NEG_INF : __ldcg(halo + ((long long)(strip - 1) * nx + q) * 8 + lane);
                    else if (NW == 1 || w == 0) hv = __ldcg(halo + (long long)q * 8 + lane);
                    else if (CS > 1 && warp == 0) hv = __ldcg(cmail + 16 + ((crank - 1) * 2 + ((t - 1) & 1)) * 8 + lane);
                    else hv = s_xbuf[warp - 1][(t - 1) & 1][lane];
                }

                float init;
                // (rounds, top end window: only paths that start in rows 0..2 count, i.e. the free starts of columns 0, 1 below
                // row 2 are closed -- see end_window_best)
                if (P.init_mode == SA_INIT_SEMIGLOBAL) init = (j == 0 || (q < 2 && (!ROUNDS || r_pre != 1 || j <= 2))) ? 0.0f : NEG_INF;
                else if (P.init_mode == SA_INIT_LOCAL) init = 0.0f;
                else init = (j == 0 || q == 0) ? 0.0f : NEG_INF;

#if SA_PRUNE
                float best = NEG_INF;
                {
                    // pass 1: upper bounds `hi` of all 36 candidates; the largest (m1, with its candidate number k6) and
                    // the second largest (m2) are tracked exactly, two candidates per update
                    float m1 = NEG_INF, m2 = NEG_INF;
                    int k6 = 0;
#pragma unroll
                    for (int dj = 1; dj <= LB; dj++) {
                        float hi[LB];
#if SA_PACKED
#pragma unroll
                        for (int c = 0; c < LB; c += 2) {  // two candidates per chain: same operations, same order, packed
                            const float sc0 = lds_f32(PC(c) - dj * 4), sc1 = lds_f32(PC(c + 1) - dj * 4);
                            float a0, a1, e0, e1, t0, t1;
                            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(a0) : "f"(fabsf(__fsub_rn(db[dj - 1], dx[c]))));
                            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(a1) : "f"(fabsf(__fsub_rn(db[dj - 1], dx[c + 1]))));
                            upk2(fma2(pk2(a0, a1), pk2(1.2f, 1.2f), pk2(-PRUNE_C, -PRUNE_C)), t0, t1);
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
                            f32x2 fnfp;  // fn_t + fp_t (SegAligner.cpp:89-91); 0 + fp_t == fp_t bit for bit (fp_t >= +0)
                            if (!SWAP) fnfp = (dj == 1) ? pk2(fx[c], fx[c + 1]) : add2(pk2(5000.0f * (float)(dj - 1), 5000.0f * (float)(dj - 1)), pk2(fx[c], fx[c + 1]));
                            else fnfp = add2(pk2(5000.0f * (float)c, 5000.0f * (float)(c + 1)), pk2(fb[dj - 1], fb[dj - 1]));
                            upk2(add2(pk2(sc0, sc1), sub2(pk2(gain, gain), add2(fnfp, pk2(e0, e1)))), hi[c], hi[c + 1]);
                        }
#else
#pragma unroll
                        for (int c = 0; c < LB; c++) {
                            const float sc = lds_f32(PC(c) - dj * 4);
                            const float e = delta_lo_approx(__fsub_rn(db[dj - 1], dx[c]));
                            float fnfp;  // fn_t + fp_t (SegAligner.cpp:89-91); 0 + fp_t == fp_t bit for bit (fp_t >= +0)
                            if (!SWAP) fnfp = (dj == 1) ? fx[c] : __fadd_rn(5000.0f * (float)(dj - 1), fx[c]);
                            else fnfp = (c == 0) ? fb[dj - 1] : __fadd_rn(5000.0f * (float)c, fb[dj - 1]);
                            hi[c] = __fadd_rn(sc, __fsub_rn(gain, __fadd_rn(fnfp, e)));
                        }
#endif
#pragma unroll
                        for (int c = 0; c < LB; c += 2) {
                            const float mx = fmaxf(hi[c], hi[c + 1]), mn = fminf(hi[c], hi[c + 1]);
                            const int kmx = (hi[c] >= hi[c + 1]) ? (dj - 1) * LB + c : (dj - 1) * LB + c + 1;
                            const float t = fminf(m1, mx);
                            k6 = (mx > m1) ? kmx : k6;
                            m1 = fmaxf(m1, mx);
                            m2 = fmaxf(m2, fmaxf(mn, t));
                        }
                    }
                    if (m1 > NEG_INF) {
                        // pass 2: the exact score of the best-looking candidate (dj, c) of this lane
                        const int djm = (k6 * 43) >> 8, c = k6 - djm * LB, dj = djm + 1;
                        const float dbv = lds_f32(a_row + djm * 128);
                        const float dxv = __ldg(reinterpret_cast<const float *>(xops + q) + c);
                        const float sc = lds_f32(PC(c) - dj * 4);
                        const float delta = powf12_exact(fabsf(__fsub_rn(dbv, dxv)), a_tab, P.pk);
                        float fnfp;
                        if (!SWAP) {
                            const float fxv = __ldg(reinterpret_cast<const float *>(xops + q) + LB + c);
                            fnfp = __fadd_rn(5000.0f * (float)djm, fxv);
                        } else {
                            const float fbv = lds_f32(a_row + (LB + djm) * 128);
                            fnfp = __fadd_rn(5000.0f * (float)c, fbv);
                        }
                        best = __fadd_rn(sc, __fsub_rn(gain, __fadd_rn(fnfp, delta)));
                        code = 1 + k6;
                        n_replay++;
                        // step 3: every other candidate's exact score is <= its hi <= m2
                        if (!(best > m2)) {
                            // rare: a tie or near-tie -> exact scan of the cell in the reference's tie order
                            const CellRes cr = cell_slow(SWAP, a_row, dx[0], dx[1], dx[2], dx[3], dx[4], dx[5],
                                                         SWAP ? 0.0f : fx[0], SWAP ? 0.0f : fx[1], SWAP ? 0.0f : fx[2],
                                                         SWAP ? 0.0f : fx[3], SWAP ? 0.0f : fx[4], SWAP ? 0.0f : fx[5],
                                                         pbase, a_tab.t2d, a_tab.exp, best, code, gain);
                            best = cr.best;
                            code = cr.code;
                            if (P.stats) atomicAdd(P.stats + 1, 1ull);
                        }
                    }
                }
#else
                float best = NEG_INF;
                uint32_t umax = 0;
                // Look-back rows dj = 1..6 as a rolled loop over row PAIRS with ping-pong operand buffers: while the
                // fp64 replay of row dj runs, the table operands of row dj+1 are already being fetched.
                PowRow rowA, rowB;
                row_prefetch(rowA, lds_f32(a_row), dx, a_tab, umax);
#pragma unroll 1
                for (int i = 0; i < LB / 2; i++) {
                    const int dj = 2 * i + 1;
                    row_prefetch(rowB, lds_f32(a_row + dj * 128), dx, a_tab, umax);  // row dj+1
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int djh = dj + h;
                        float delta[LB], sc[LB], ns[LB];
#pragma unroll
                        for (int c = 0; c < LB; c++) {
                            sc[c] = lds_f32(PC(c) - djh * 4);
                        }
                        if (h == 0) {
                            row_powf(rowA, a_tab, P.pk, delta);
                            if (i < LB / 2 - 1) row_prefetch(rowA, lds_f32(a_row + (dj + 1) * 128), dx, a_tab, umax);  // row dj+2
                        } else {
                            row_powf(rowB, a_tab, P.pk, delta);
                        }
                        // fn_t + fp_t (SegAligner.cpp:89-91); with swap_b_x the maps exchange roles (:119).
                        // fn_t = 5000*(dj-1) resp. 5000*(dq-1) is exact; 0 + fp_t == fp_t bit for bit (fp_t >= +0).
                        const float fn = 5000.0f * (float)(djh - 1);
                        float fbv = 0.0f;
                        if (SWAP) fbv = lds_f32(a_row + (LB + djh - 1) * 128);
#pragma unroll
                        for (int c = 0; c < LB; c++) {
                            float fnfp;
                            if (!SWAP) fnfp = __fadd_rn(fn, fx[c]);
                            else fnfp = (c == 0) ? fbv : __fadd_rn(5000.0f * (float)c, fbv);
                            const float r = __fsub_rn(10000.0f, __fadd_rn(fnfp, delta[c]));
                            ns[c] = __fadd_rn(sc[c], r);
                        }
                        if (STORE) {  // last maximum in scan order (dj ascending, dq ascending) == first in the reference's
                            float rb = ns[0];
                            int rc = 0;
#pragma unroll
                            for (int c = 1; c < LB; c++)
                                if (ns[c] >= rb) {
                                    rb = ns[c];
                                    rc = c;
                                }
                            if (rb >= best) {
                                best = rb;
                                code = 1 + (djh - 1) * LB + rc;
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < LB; c++) best = fmaxf(best, ns[c]);
                        }
                    }
                }
                if (umax >= IX_RANGE) {
                    // rare: some discrepancy of this cell is exactly 0, tiny, huge, inf or NaN, where the branch-free
                    // replay is undefined -> redo the whole cell, candidate by candidate
                    const CellRes cr = cell_slow(SWAP, a_row, dx[0], dx[1], dx[2], dx[3], dx[4], dx[5], SWAP ? 0.0f : fx[0],
                                                 SWAP ? 0.0f : fx[1], SWAP ? 0.0f : fx[2], SWAP ? 0.0f : fx[3],
                                                 SWAP ? 0.0f : fx[4], SWAP ? 0.0f : fx[5], pbase, a_tab.t2d, a_tab.exp,
                                                 NEG_INF, 0);
                    best = cr.best;
                    code = cr.code;
                }
#endif
                val = init;
                if (best > init && !row_used) val = best;
                else code = 0;
                rv = val;
                if (P.fitting && j == 0 && q > 0) rv = NEG_INF;  // dp_align's `i_ind == 0 && p_ind > 0` break (:111)
            }
            __syncwarp();
            if (active) {
                // publish column q: own row, the 6 rows above the strip, and our bottom rows for the strip below
                sts_f32(pbase, rv);
                sts_f32(pbase + LB * RING_W * 4, rv);
#ifdef SA_CPL_DEBUG
                if (CPL && lane == 0) *reinterpret_cast<volatile int32_t *>(P.cpl_done + P.n_problems + 64 + cslot + cu) = 0x40000000 | (strip << 12) | q;
#endif
                if (CPL && strip > 0 && lane < 6) {
                    // the entry loaded at the start of the step is valid once it carries the tag of (strip - 1, q)
                    const uint32_t want = P.cpl_epoch | ((uint32_t)(strip - 1) * (uint32_t)nx + (uint32_t)q + 1u);
                    for (unsigned spin = 0; (uint32_t)(hv_cpl >> 32) != want; spin++) {
                        // cpl_spin_limit polls (>= ~6 s, ten times the longest unit of the launch) without the producer's entry: it is not running.  Fail loudly on the host; once the
                        // flag is up every other wait gives up at once, so the kernel drains instead of hanging.
                        if (spin > P.cpl_spin_limit || ((spin & 1023u) == 1023u && *reinterpret_cast<volatile int32_t *>(P.cpl_error) != 0)) {
                            if (atomicCAS(P.cpl_error, 0, 1) == 0) {  // first failure: leave a description for the host's error text
                                P.cpl_error[1] = (int32_t)pidx, P.cpl_error[2] = cu, P.cpl_error[3] = cU, P.cpl_error[4] = strip;
                                P.cpl_error[5] = q, P.cpl_error[6] = (int32_t)want, P.cpl_error[7] = (int32_t)(hv_cpl >> 32);
                                P.cpl_error[8] = nb, P.cpl_error[9] = nx, P.cpl_error[10] = (int32_t)t_idx, P.cpl_error[11] = lane;
#ifdef SA_CPL_DEBUG
                                for (int k = 0; k < cU && k < 16; k++)
                                    P.cpl_error[16 + k] = *reinterpret_cast<volatile int32_t *>(P.cpl_done + P.n_problems + 64 + cslot + k);
                                P.cpl_error[12] = (int32_t)*reinterpret_cast<volatile unsigned long long *>(P.counter);
#endif
                            }
                            break;
                        }
                        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(hv_cpl) : "l"(mb_in + (long long)q * 6 + lane) : "memory");
                    }
                    hv = __uint_as_float((uint32_t)hv_cpl);
                }
                if (lane < 6) {
                    sts_f32(pbase - 24, hv);  // a_halo = a_my - 6 floats
                    sts_f32(pbase - 24 + LB * RING_W * 4, hv);
                }
                // Coupled units: this column may only be handed on AFTER the column above it has been taken over.  That order is
                // what keeps a producer from overwriting a mailbox entry its consumer still needs (the ring argument above), and
                // divergent lanes of a warp are not ordered by themselves: measured on B200, lanes 26..31 published while lanes
                // 0..5 were still polling, a whole ring ran ahead on it and overwrote entries (hand-overs then never arrived).
                if (CPL) __syncwarp();
                if (lane >= 26) {
                    if (CPL) {
                        const unsigned long long e = ((unsigned long long)(P.cpl_epoch | ((uint32_t)strip * (uint32_t)nx + (uint32_t)q + 1u)) << 32) | __float_as_uint(rv);
                        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mb_out + (long long)q * 6 + (lane - 26)), "l"(e) : "memory");
                    } else if (ROUNDS) __stcg(halo + ((long long)strip * nx + q) * 8 + (lane - 26), rv);
                    else if (NW == 1 || w == NWT - 1) __stcg(halo + (long long)q * 8 + (lane - 26), rv);
                    else if (CS > 1 && warp == NW - 1) __stcg(cmail + 16 + (crank * 2 + (t & 1)) * 8 + (lane - 26), rv);
                    else s_xbuf[warp][t & 1][lane - 26] = rv;
                }
            }
            if (NW == 1) __syncwarp();
            else if (CS == 1) __syncthreads();
            else cg::this_cluster().sync();
            if (active) {
                if (STORE) {
                    // stage this column; every 8 columns (and at the end of the row) flush the 32 x 8 tile: 8 lanes
                    // write the 8 consecutive columns of one row, 4 rows per instruction
                    const int cq = q & 7;
                    if (STORE_S) sts_f32(a_tile_s + lane * 36 + cq * 4, val);
                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a_tile_c + lane * 12 + cq), "r"(code) : "memory");
                    if (cq == 7 || q == nx - 1) {
                        __syncwarp();
                        long long coff;
                        asm volatile("ld.shared.s64 %0, [%1];" : "=l"(coff) : "r"(a_tile_c + 32 * 12));
                        float *Sout = STORE_S ? P.S + coff : nullptr;
                        uint8_t *Cout = (ROUNDS ? P.r_code : P.code) + coff;
                        const int q0 = q - cq, col = lane & 7;
#pragma unroll
                        for (int it = 0; it < 8; it++) {
                            const int r = it * 4 + (lane >> 3);
                            const int jr = (strip << 5) + r;
                            if (col <= cq && jr < nb) {
                                const long long o = (long long)jr * nx + q0 + col;
                                uint32_t cv;
                                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(cv) : "r"(a_tile_c + r * 12 + col));
                                if (STORE_S) Sout[o] = lds_f32(a_tile_s + r * 36 + col * 4);
                                Cout[o] = (uint8_t)cv;
                            }
                        }
                        __syncwarp();
                    }
                }
                if (row_ok) {
                    if (STORE_S) {
                        if (P.rowmax_s) {
                            if (val > 0.0f && val >= rm_s) {
                                rm_s = val;
                                rm_q = q;
                            }
                            if (q == nx - 1) {
                                P.rowmax_s[P.row_off[pidx] + j] = rm_s;
                                P.rowmax_q[P.row_off[pidx] + j] = rm_q;
                            }
                        }
                    }
                    if (ROUNDS) {  // what the backtrack-start scan reads persists per pair (rounds_finish)
                        if (q >= nx - 2) __stcg(P.r_col2 + gunit * 2 * P.r_max_nb + 2 * j + (q - (nx - 2)), val);
                        if (j == nb - 1) __stcg(P.r_lastrow + gunit * P.r_max_nx + q, val);
                    } else if (P.start_mode == SA_START_LOCAL) {
                        if (val > st_s) {  // per lane (j,q) arrives in increasing row-major order
                            st_s = val;
                            st_j = j;
                            st_q = q;
                        }
                    } else if (P.start_mode == SA_START_SEMIGLOBAL && (q >= nx - 2 || strip == nstrips - 1)) {  // warp-uniform
                        // scan order of sg_aln_backtrack_start: last row by q, then rows by j over the last 2 columns
                        const bool last_row = (j == nb - 1);
                        const bool last_cols = (q >= nx - 2);
                        if (last_row || last_cols) {
                            if (val > st_s || (val == st_s && sg_ord(j, q) < sg_ord(st_j, st_q))) {
                                st_s = val;
                                st_j = j;
                                st_q = q;
                            }
                        }
                    }
                }
                slot = (slot == LB - 1) ? 0 : slot + 1;
                pbase = a_my + slot * (RING_W * 4);
            }
            if (NW == 1 || (t >= w && strip < nstrips)) {  // advance within the round (idle steps of a short round included)
                q++;
                if (q == R) {
                    q = 0;
                    strip += CPL ? cU : NWT;
                }
            }
        }
        if (!ROUNDS) break;
        __syncwarp();
        if (r_pre == 1) {  // top strips done: now the bottom strips, as if nothing lay above them
            r_pre = 2;
            r_sb = (nb - 4 - win_span(P, nx)) >> 5;
            r_se = nstrips;
            continue;
        }
        if (r_pre == 2) {
            if (end_window_best(P, gunit, nb, nx) < P.r_win_min) {  // no round of this pair can write a tip alignment
                if (lane == 0) P.r_n_rounds[pidx] = -2;
                break;
            }
            r_pre = 0;  // round 1: the whole matrix (the top strips were swept with the starts below row 2 closed)
            r_sb = 0;
            r_se = nstrips;
            continue;
        }
        if (rounds_finish(P, gunit, pidx, pr.contig, nb, pr.seg, nx, P.r_thr[pidx], r_round, r_count, r_sb, r_se)) break;
        r_round++;
        }  // rounds
        if (ROUNDS) {
            if (P.stats) {
                const unsigned a = __reduce_add_sync(0xffffffffu, n_replay);
                if (lane == 0) atomicAdd(P.stats, (unsigned long long)a);
            }
            continue;
        }

        // reduce the start over lanes (and warps): max score, ties -> smallest scan order
        if (P.start_mode == SA_START_LOCAL) st_ord = (st_j < 0) ? 0x7fffffffffffffffll : (long long)st_j * nx + st_q;
        else if (P.start_mode == SA_START_SEMIGLOBAL) st_ord = sg_ord(st_j, st_q);
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, st_s, m);
            const int oj = __shfl_xor_sync(0xffffffffu, st_j, m);
            const int oq = __shfl_xor_sync(0xffffffffu, st_q, m);
            const long long oo = __shfl_xor_sync(0xffffffffu, st_ord, m);
            const bool take = (oj >= 0) && (st_j < 0 || os > st_s || (os == st_s && oo < st_ord));
            if (take) {
                st_s = os;
                st_j = oj;
                st_q = oq;
                st_ord = oo;
            }
        }
        if (NW > 1) {
            if (lane == 0) {
                s_red_s[warp] = st_s;
                s_red_j[warp] = st_j;
                s_red_q[warp] = st_q;
                s_red_o[warp] = st_ord;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int k = 1; k < NW; k++) {
                    const bool take = (s_red_j[k] >= 0) &&
                                      (st_j < 0 || s_red_s[k] > st_s || (s_red_s[k] == st_s && s_red_o[k] < st_ord));
                    if (take) {
                        st_s = s_red_s[k];
                        st_j = s_red_j[k];
                        st_q = s_red_q[k];
                        st_ord = s_red_o[k];
                    }
                }
            }
            if (CS > 1) {  // per-CTA candidates through the mailbox, CTA 0 picks the winner
                float *slot = cmail + 16 + CS * 16 + crank * 8;
                if (threadIdx.x == 0) {
                    __stcg(slot + 0, st_s);
                    __stcg(reinterpret_cast<int *>(slot) + 1, st_j);
                    __stcg(reinterpret_cast<int *>(slot) + 2, st_q);
                    __stcg(reinterpret_cast<long long *>(slot) + 2, st_ord);
                }
                cg::this_cluster().sync();
                if (crank == 0 && threadIdx.x == 0) {
                    for (int k = 1; k < CS; k++) {
                        const float *o = cmail + 16 + CS * 16 + k * 8;
                        const float os = __ldcg(o);
                        const int oj = __ldcg(reinterpret_cast<const int *>(o) + 1);
                        const int oq = __ldcg(reinterpret_cast<const int *>(o) + 2);
                        const long long oo = __ldcg(reinterpret_cast<const long long *>(o) + 2);
                        if (oj >= 0 && (st_j < 0 || os > st_s || (os == st_s && oo < st_ord))) {
                            st_s = os;
                            st_j = oj;
                            st_q = oq;
                            st_ord = oo;
                        }
                    }
                }
            }
        }
        if (P.stats) {
            const unsigned a = __reduce_add_sync(0xffffffffu, n_replay);
            if (lane == 0) atomicAdd(P.stats, (unsigned long long)a);
        }
        if (CPL) {
            // every unit leaves its candidate; the unit that finishes last picks the winner (same rule as between lanes)
            bool last = false;
            if (lane == 0) {
                CplRes r;
                r.score = st_s, r.j = st_j, r.q = st_q, r.pad = 0, r.ord = st_ord;
                P.cpl_res[cslot + cu] = r;
                __threadfence();
                last = atomicAdd(P.cpl_done + cslot, 1) == cU - 1;
                if (last) {
                    __threadfence();
                    st_j = -1;
                    for (int k = 0; k < cU; k++) {
                        const CplRes *o = P.cpl_res + cslot + k;
                        const float os = __ldcg(&o->score);
                        const int oj = __ldcg(&o->j), oq = __ldcg(&o->q);
                        const long long oo = __ldcg(&o->ord);
                        if (oj >= 0 && (st_j < 0 || os > st_s || (os == st_s && oo < st_ord))) {
                            st_s = os;
                            st_j = oj;
                            st_q = oq;
                            st_ord = oo;
                        }
                    }
                }
            }
            if (!__shfl_sync(0xffffffffu, (int)last, 0)) continue;  // the result is written by lane 0 of the last unit
        }
        if ((NW == 1 && lane == 0) || (NW > 1 && threadIdx.x == 0 && crank == 0)) {
            sa_result res;
            if (P.start_mode == SA_START_END) {
                res.j = nb - 1;
                res.q = nx - 1;
                res.score = 0.0f;  // filled by the traceback kernel from the stored matrix
            } else {
                res.j = st_j;
                res.q = st_q;
                res.score = st_s;
            }
            res.path_len = 0;
            P.results[pidx] = res;
        }
    }
}

// -nl: dp_align with lookback = 9999999 (main.cpp:475-478): every i < j and p < q is a predecessor, O(nb^2 nx^2)
// ("not recommended for large datasets", README.md:222).  Straight restatement, one CTA per problem: rows are
// sequential, threads take the columns of a row; each thread scans its predecessors in the reference's order
// (i ascending, p ascending, strict '>') with the generic powf path and the sequential fp32 accumulate of score_f.
// S and the (i,p) predecessor (packed i<<16|p, -1 = none) are always stored; the start scan runs afterwards.
__global__ void __launch_bounds__(128) fill_unbanded_kernel(const FillParams P, int32_t *prev32, int swap) {
    __shared__ float sh_s[128];
    __shared__ int sh_j[128], sh_q[128];
    __shared__ long long sh_o[128];
    for (long long t = blockIdx.x; t < P.n_problems; t += gridDim.x) {
        const long long pidx = P.order ? P.order[t] : t;
        const sa_problem pr = P.problems[pidx];
        const int64_t boff = P.contigs.pos_off[pr.contig];
        const int nb = (int)(P.contigs.pos_off[pr.contig + 1] - boff) - 1;
        const float *b = P.contigs.pos + boff;
        const float *bprob = P.contigs.prob + (boff - pr.contig);
        const int64_t xoff = P.segs.pos_off[pr.seg];
        const int nx = (int)(P.segs.pos_off[pr.seg + 1] - xoff) - 1;
        const float *x = P.segs.pos + xoff;
        const float *xprob = P.segs.prob + (xoff - pr.seg);
        const uint32_t *used =
            (P.used_bits && pr.used_slot >= 0) ? P.used_bits + (long long)pr.used_slot * P.used_words_per_slot : nullptr;
        float *S = P.S + P.cell_off[pidx];
        int32_t *PV = prev32 + P.cell_off[pidx];
        for (int j = 0; j < nb; j++) {
            const bool row_used = used ? ((used[j >> 5] >> (j & 31)) & 1u) : false;
            float rm_s = 0.0f;
            int rm_q = -1;
            for (int q = threadIdx.x; q < nx; q += blockDim.x) {
                float cur;
                if (P.init_mode == SA_INIT_SEMIGLOBAL) cur = (j == 0 || q < 2) ? 0.0f : NEG_INF;
                else if (P.init_mode == SA_INIT_LOCAL) cur = 0.0f;
                else cur = (j == 0 || q == 0) ? 0.0f : NEG_INF;
                int pv = -1;
                if (!row_used) {
                    for (int i = 0; i < j; i++) {
                        for (int p = 0; p < q; p++) {
                            if (P.fitting && i == 0 && p > 0) break;
                            float acc = 0.0f, fn;
                            if (!swap) {
                                for (int k = p + 1; k < q; k++) acc = __fadd_rn(acc, xprob[k]);
                                fn = 5000.0f * (float)(j - (i + 1));
                            } else {
                                for (int k = i + 1; k < j; k++) acc = __fadd_rn(acc, bprob[k]);
                                fn = 5000.0f * (float)(q - (p + 1));
                            }
                            const float discrep = fabsf(__fsub_rn(__fsub_rn(b[j], b[i]), __fsub_rn(x[q], x[p])));
                            const float delta = powf12_slow(discrep);
                            const float fp = __fmul_rn(5000.0f, acc);
                            const float sc = __fsub_rn(10000.0f, __fadd_rn(__fadd_rn(fn, fp), delta));
                            const float ns = __fadd_rn(S[(long long)i * nx + p], sc);
                            if (ns > cur) {
                                cur = ns;
                                pv = (i << 16) | p;
                            }
                        }
                    }
                }
                S[(long long)j * nx + q] = cur;
                PV[(long long)j * nx + q] = pv;
                if (cur > 0.0f && cur >= rm_s) {
                    rm_s = cur;
                    rm_q = q;
                }
            }
            if (P.rowmax_s) {  // per-row maximum for the detection harvest (ties -> larger q)
                sh_s[threadIdx.x] = rm_s;
                sh_q[threadIdx.x] = rm_q;
                __syncthreads();
                for (int m = 64; m > 0; m >>= 1) {
                    if ((int)threadIdx.x < m) {
                        const float os = sh_s[threadIdx.x + m];
                        const int oq = sh_q[threadIdx.x + m];
                        if (oq >= 0 && (sh_q[threadIdx.x] < 0 || os > sh_s[threadIdx.x] ||
                                        (os == sh_s[threadIdx.x] && oq > sh_q[threadIdx.x]))) {
                            sh_s[threadIdx.x] = os;
                            sh_q[threadIdx.x] = oq;
                        }
                    }
                    __syncthreads();
                }
                if (threadIdx.x == 0) {
                    P.rowmax_s[P.row_off[pidx] + j] = sh_q[0] >= 0 ? sh_s[0] : 0.0f;
                    P.rowmax_q[P.row_off[pidx] + j] = sh_q[0];
                }
            }
            __syncthreads();  // row j complete (global writes visible to the CTA) before row j+1 reads it
        }
        // backtrack start: scan in the reference's order (see fill_kernel), parallel max with order keys
        float st_s = NEG_INF;
        int st_j = -1, st_q = -1;
        long long st_ord = 0x7fffffffffffffffll;
        if (P.start_mode != SA_START_END) {
            for (long long c = threadIdx.x; c < (long long)nb * nx; c += blockDim.x) {
                const int j = (int)(c / nx), q = (int)(c % nx);
                long long ord;
                if (P.start_mode == SA_START_LOCAL) ord = c;
                else {
                    const bool last_row = (j == nb - 1), last_cols = (q >= nx - 2);
                    if (!last_row && !last_cols) continue;
                    ord = last_row ? (long long)q : (long long)nx + 2ll * j + (q - (nx - 2));
                }
                const float v = S[c];
                if (v > st_s || (v == st_s && ord < st_ord && v > NEG_INF)) {
                    st_s = v;
                    st_j = j;
                    st_q = q;
                    st_ord = ord;
                }
            }
            sh_s[threadIdx.x] = st_s;
            sh_j[threadIdx.x] = st_j;
            sh_q[threadIdx.x] = st_q;
            sh_o[threadIdx.x] = st_ord;
            __syncthreads();
            for (int m = 64; m > 0; m >>= 1) {
                if ((int)threadIdx.x < m) {
                    const int o = threadIdx.x + m;
                    if (sh_j[o] >= 0 && (sh_j[threadIdx.x] < 0 || sh_s[o] > sh_s[threadIdx.x] ||
                                         (sh_s[o] == sh_s[threadIdx.x] && sh_o[o] < sh_o[threadIdx.x]))) {
                        sh_s[threadIdx.x] = sh_s[o];
                        sh_j[threadIdx.x] = sh_j[o];
                        sh_q[threadIdx.x] = sh_q[o];
                        sh_o[threadIdx.x] = sh_o[o];
                    }
                }
                __syncthreads();
            }
        }
        if (threadIdx.x == 0) {
            sa_result res;
            if (P.start_mode == SA_START_END) {
                res.j = nb - 1;
                res.q = nx - 1;
                res.score = S[(long long)(nb - 1) * nx + (nx - 1)];
            } else {
                res.j = sh_j[0];
                res.q = sh_q[0];
                res.score = sh_s[0];
            }
            res.path_len = 0;
            P.results[pidx] = res;
        }
        __syncthreads();
    }
}

// get_aln_list (SAHelper.cpp:83-95): follow the stored predecessor codes from the start cell.
__global__ void trace_kernel(const TraceParams P) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P.n_problems) return;
    const sa_problem pr = P.problems[k];
    const int64_t xoff = P.segs.pos_off[pr.seg];
    const int nx = (int)(P.segs.pos_off[pr.seg + 1] - xoff) - 1;
    // S == nullptr: only the predecessor codes were stored; the scores along the path are recomputed below
    const float *S = P.S ? P.S + P.cell_off[k] : nullptr;
    const uint8_t *C = P.code ? P.code + P.cell_off[k] : nullptr;
    sa_result res = P.results[k];
    int j = res.j, q = res.q;
    // count_only: no output arrays yet, the capacity is the longest possible path min(nb, nx)
    const long long base = P.count_only ? 0 : P.path_off[k];
    long long cap;
    if (P.count_only) {
        const int nb_ = (int)(P.contigs.pos_off[pr.contig + 1] - P.contigs.pos_off[pr.contig]) - 1;
        cap = nb_ < nx ? nb_ : nx;
    } else {
        cap = P.path_off[k + 1] - base;
    }
    int n = 0;
    if (j >= 0 && S) res.score = S[(long long)j * nx + q];
    while (j >= 0 && n < cap) {
        const long long c = (long long)j * nx + q;
        if (!P.count_only) {
            P.path_j[base + n] = j;
            P.path_q[base + n] = q;
            if (S) P.path_s[base + n] = S[c];
        }
        n++;
        if (P.prev32) {
            const int pv = P.prev32[P.cell_off[k] + c];
            if (pv < 0) break;
            j = pv >> 16;
            q = pv & 0xffff;
        } else {
            const int cd = C[c];
            if (cd == 0) break;
            const int dj = (cd - 1) / LB + 1, dq = (cd - 1) % LB + 1;
            j -= dj;
            q -= dq;
        }
    }
    if (!S && !P.count_only && n > 0) {
        // Scores along the path, recomputed exactly as dp_align produced them (SegAligner.cpp:119-127): the path's first
        // cell kept its initial value (code 0), every later cell is S[pred] + score_f(pred -> cell) in the reference's
        // fp32 operation order with the bit-exact powf.  Saves storing 4 of the 5 bytes per DP cell.
        const int64_t boff = P.contigs.pos_off[pr.contig];
        const LabelOps *bops = P.contigs.ops + (boff - pr.contig);
        const LabelOps *xops = P.segs.ops + (xoff - pr.seg);
        const int j0 = P.path_j[base + n - 1], q0 = P.path_q[base + n - 1];
        float sv;
        if (P.init_mode == SA_INIT_SEMIGLOBAL) sv = (j0 == 0 || q0 < 2) ? 0.0f : NEG_INF;
        else if (P.init_mode == SA_INIT_LOCAL) sv = 0.0f;
        else sv = (j0 == 0 || q0 == 0) ? 0.0f : NEG_INF;
        P.path_s[base + n - 1] = sv;
        int jp = j0, qp = q0;
        for (int t = n - 2; t >= 0; t--) {
            const int jc = P.path_j[base + t], qc = P.path_q[base + t];
            const int dj = jc - jp, dq = qc - qp;
            const float discrep = fabsf(__fsub_rn(bops[jc].d[dj - 1], xops[qc].d[dq - 1]));
            const float delta = powf12_slow(discrep);
            float fnfp;
            if (!P.swap) fnfp = (dj == 1) ? xops[qc].f[dq - 1] : __fadd_rn(5000.0f * (float)(dj - 1), xops[qc].f[dq - 1]);
            else fnfp = (dq == 1) ? bops[jc].f[dj - 1] : __fadd_rn(5000.0f * (float)(dq - 1), bops[jc].f[dj - 1]);
            sv = __fadd_rn(sv, __fsub_rn(10000.0f, __fadd_rn(fnfp, delta)));
            P.path_s[base + t] = sv;
            jp = jc;
            qp = qc;
        }
        res.score = sv;
    }
    res.path_len = n;
    P.results[k] = res;
    if (P.len_out) P.len_out[k] = n;
}

// multi_backtrack_scores (SegAligner.cpp:208-233) + zero_out_aln (:198-205) for one problem per CTA.
// The reference sorts every cell with S > 0 ascending by (score, j, q), walks the list from the top, ALWAYS
// skips the top entry (:224 decrements before use), and takes an entry when its current S is still > 0, then
// sets its whole traced path to -inf.  Equivalent formulation used here: kill the top cell, then repeatedly
// take the largest (score, j, q) among cells that are still > 0 and wipe its path.  Per-row maxima (filled by
// the sweep) make each selection an O(nb) reduction plus an O(nx) re-scan of the rows the wiped path touched.
__global__ void __launch_bounds__(256) detect_kernel(const DetectParams P) {
    const long long k = blockIdx.x;
    if (k >= P.n_problems) return;
    const sa_problem pr = P.problems[k];
    const int64_t boff = P.contigs.pos_off[pr.contig];
    const int nb = (int)(P.contigs.pos_off[pr.contig + 1] - boff) - 1;
    const int64_t xoff = P.segs.pos_off[pr.seg];
    const int nx = (int)(P.segs.pos_off[pr.seg + 1] - xoff) - 1;
    float *S = P.S + P.cell_off[k];
    const uint8_t *C = P.code ? P.code + P.cell_off[k] : nullptr;
    const int32_t *PV = P.prev32 ? P.prev32 + P.cell_off[k] : nullptr;
    float *rms = P.rowmax_s + P.row_off[k];
    int *rmq = P.rowmax_q + P.row_off[k];
    const int want = P.n_want[k];
    float *out = P.scores_out + P.scores_off[k];
    __shared__ float sh_s[256];
    __shared__ int sh_j[256], sh_q[256];
    __shared__ int sh_bj, sh_bq;
    __shared__ float sh_bs;
    const int tid = threadIdx.x;
    int got = 0;
    for (int it = 0; it <= want; it++) {
        // argmax over rows: score desc, then j desc, then q desc
        float bs = 0.0f;
        int bj = -1, bq = -1;
        for (int r = tid; r < nb; r += 256) {
            const float s = rms[r];
            if (s > 0.0f && (s > bs || (s == bs && r > bj))) {
                bs = s;
                bj = r;
                bq = rmq[r];
            }
        }
        sh_s[tid] = bs;
        sh_j[tid] = bj;
        sh_q[tid] = bq;
        __syncthreads();
        for (int m = 128; m > 0; m >>= 1) {
            if (tid < m) {
                const float os = sh_s[tid + m];
                const int oj = sh_j[tid + m];
                if (oj >= 0 && (sh_j[tid] < 0 || os > sh_s[tid] || (os == sh_s[tid] && oj > sh_j[tid]))) {
                    sh_s[tid] = os;
                    sh_j[tid] = oj;
                    sh_q[tid] = sh_q[tid + m];
                }
            }
            __syncthreads();
        }
        if (tid == 0) {
            sh_bs = sh_s[0];
            sh_bj = sh_j[0];
            sh_bq = sh_q[0];
        }
        __syncthreads();
        if (sh_bj < 0) break;  // no positive cell left (the reference would index all_scores[-1] here)
        if (it > 0) {
            if (tid == 0) out[got] = sh_bs;
            got++;
        }
        // wipe: the top cell alone on the first pass (it is only skipped, never traced), else the whole path
        int j = sh_bj, q = sh_bq;
        __syncthreads();
        for (;;) {
            const long long c = (long long)j * nx + q;
            int nj = -1, nq = -1;  // predecessor on the traced path
            if (PV) {
                const int pv = PV[c];
                if (pv >= 0) {
                    nj = pv >> 16;
                    nq = pv & 0xffff;
                }
            } else {
                const int cd = C[c];
                if (cd != 0) {
                    nj = j - ((cd - 1) / LB + 1);
                    nq = q - ((cd - 1) % LB + 1);
                }
            }
            __syncthreads();
            if (tid == 0) S[c] = NEG_INF;
            __syncthreads();
            // re-scan row j
            float rs = 0.0f;
            int rq = -1;
            for (int x = tid; x < nx; x += 256) {
                const float s = S[(long long)j * nx + x];
                if (s > 0.0f && s >= rs) {
                    rs = s;
                    rq = x;
                }
            }
            sh_s[tid] = rs;
            sh_q[tid] = rq;
            __syncthreads();
            for (int m = 128; m > 0; m >>= 1) {
                if (tid < m) {
                    const float os = sh_s[tid + m];
                    const int oq = sh_q[tid + m];
                    if (oq >= 0 && (sh_q[tid] < 0 || os > sh_s[tid] || (os == sh_s[tid] && oq > sh_q[tid]))) {
                        sh_s[tid] = os;
                        sh_q[tid] = oq;
                    }
                }
                __syncthreads();
            }
            if (tid == 0) {
                rms[j] = sh_q[0] >= 0 ? sh_s[0] : 0.0f;
                rmq[j] = sh_q[0];
            }
            __syncthreads();
            if (it == 0 || nj < 0) break;
            j = nj;
            q = nq;
        }
    }
    if (tid == 0) P.n_got[k] = got;
}

__global__ void powf12_kernel(const float *x, long long n, float *out, int fast) {
    extern __shared__ unsigned char s_dyn[];
    PowTabSmem &s_tab = *tab_place(s_dyn);
    powtab_init(s_tab);
    __syncthreads();
    const TabAddr a_tab = tab_addr(s_tab);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float ax = fabsf(x[i]);
        uint32_t u;
        const float f = powf12_fast(ax, a_tab, c_pk, u);
        out[i] = (fast && u < IX_RANGE) ? f : powf12_slow(ax);
    }
}

__global__ void powf12_range_kernel(uint32_t lo, uint32_t hi, float *out_fast, float *out_generic) {
    extern __shared__ unsigned char s_dyn[];
    PowTabSmem &s_tab = *tab_place(s_dyn);
    powtab_init(s_tab);
    __syncthreads();
    const TabAddr a_tab = tab_addr(s_tab);
    const unsigned long long n = (unsigned long long)hi - lo;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float ax = __uint_as_float(lo + (uint32_t)i);
        uint32_t u;
        const float f = powf12_fast(ax, a_tab, c_pk, u);
        const float g = powf12_slow(ax);
        out_fast[i] = (u >= IX_RANGE) ? g : f;
        out_generic[i] = g;
    }
}

// Certification of the pruning bound: counts the bit patterns x in [lo, hi) for which delta_lo_approx(x) is NOT a lower
// bound of the bit-exact powf(x, 1.2f) (must be 0 for the pruning to be exact), and returns the smallest slack seen.
__global__ void prune_bound_kernel(uint32_t lo, uint32_t hi, unsigned long long *n_viol, float *min_ratio) {
    extern __shared__ unsigned char s_dyn[];
    PowTabSmem &s_tab = *tab_place(s_dyn);
    powtab_init(s_tab);
    __syncthreads();
    const TabAddr a_tab = tab_addr(s_tab);
    const unsigned long long n = (unsigned long long)hi - lo;
    unsigned long long viol = 0;
    float mr = 2.0f;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float ax = __uint_as_float(lo + (uint32_t)i);
        const float exact = powf12_exact(ax, a_tab, c_pk);
        const float lob = delta_lo_approx(ax);
        const float lob_neg = delta_lo_approx(-ax);  // the kernel feeds the signed difference; |.| is inside
        if (!(lob <= exact) || lob != lob_neg) viol++;
        if (exact > 0.0f && exact < CUDART_INF_F) mr = fminf(mr, exact / lob);
    }
    if (viol) atomicAdd(n_viol, viol);
    // positive floats order like their bit patterns
    atomicMin(reinterpret_cast<int *>(min_ratio), __float_as_int(mr));
}

// Sustained DFMA issue rate: 8 independent chains per thread, 1024 threads per SM.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *sink, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = __fma_rn(a0, m, c);
            a1 = __fma_rn(a1, m, c);
            a2 = __fma_rn(a2, m, c);
            a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c);
            a5 = __fma_rn(a5, m, c);
            a6 = __fma_rn(a6, m, c);
            a7 = __fma_rn(a7, m, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

// Sustained MUFU (XU pipe) issue rate: 8 independent lg2/ex2 chains per thread, 1024 threads per SM.
__global__ void __launch_bounds__(256) mufu_peak_kernel(float *sink, int iters) {
    float a[8];
#pragma unroll
    for (int u = 0; u < 8; u++) a[u] = 1.5f + threadIdx.x * 1e-3f + u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int u = 0; u < 8; u++) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[u]));
#pragma unroll
            for (int u = 0; u < 8; u++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[u]));
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int u = 0; u < 8; u++) s += a[u];
    if (s == 123.456f) sink[0] = s;  // never true; keeps the chains alive
}

// dynamic shared memory of a fill CTA with `wpc` warps: table region + mirrored ring + row operands
static size_t fill_dyn_smem(int wpc, bool store = true) {
    return TAB_SMEM_BYTES + (size_t)wpc * (2 * LB * RING_W + 2 * LB * 32) * sizeof(float) +
           (store ? (size_t)wpc * STAGE_BYTES_PER_WARP : 0);
}
template <int NW, int MINB, int CS, bool CPL = false>
static void launch_fill_t(const FillParams &p, int store, bool swap, int n_ctas, cudaStream_t st) {
    const size_t sm = fill_dyn_smem(NW == 1 ? FILL_WARPS : NW, store != 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_ctas);
    cfg.blockDim = dim3((NW == 1 ? FILL_WARPS : NW) * 32);
    cfg.dynamicSmemBytes = sm;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (CS > 1) ? 1 : 0;
#define SA_LAUNCH_FILL(S_, W_)                                                                                         \
    do {                                                                                                               \
        cudaFuncSetAttribute(fill_kernel<S_, W_, NW, MINB, CS, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        cudaLaunchKernelEx(&cfg, fill_kernel<S_, W_, NW, MINB, CS, CPL>, p);                                                \
    } while (0)
    if (store == 3) {  // the rounds variant exists for the bulk shape only (one warp per pair), semi-global, no swap
        if constexpr (NW == 1 && CS == 1 && !CPL) SA_LAUNCH_FILL(3, false);
    } else if (store == 2) {
        if (swap) SA_LAUNCH_FILL(2, true);
        else SA_LAUNCH_FILL(2, false);
    } else if (store == 1) {
        if (swap) SA_LAUNCH_FILL(1, true);
        else SA_LAUNCH_FILL(1, false);
    } else {
        if (swap) SA_LAUNCH_FILL(0, true);
        else SA_LAUNCH_FILL(0, false);
    }
#undef SA_LAUNCH_FILL
}
// shape: 0 = bulk (one warp per problem), 1 = cooperative CTA, 2 = cooperative cluster of CLUSTER_CTAS CTAs
void launch_fill(const FillParams &p, int store, bool swap, int shape, int n_ctas, cudaStream_t st) {
    if (shape == 3) launch_fill_t<1, FILL_MINB, 1, true>(p, store, swap, n_ctas, st);
    else if (shape == 2) launch_fill_t<CLUSTER_WARPS, 1, CLUSTER_CTAS>(p, store, swap, n_ctas, st);
    else if (shape == 1) launch_fill_t<COOP_WARPS, COOP_MINB, 1>(p, store, swap, n_ctas, st);
    else launch_fill_t<1, FILL_MINB, 1>(p, store, swap, n_ctas, st);
}

void launch_fill_unbanded(const FillParams &p, int32_t *prev32, bool swap, int n_ctas, cudaStream_t st) {
    fill_unbanded_kernel<<<n_ctas, 128, 0, st>>>(p, prev32, swap ? 1 : 0);
}

void fill_powf_constants(double *pk) {
    const double v[12] = {SA_POWF_A0, SA_POWF_A1, SA_POWF_A2, SA_POWF_A3, SA_POWF_A4, SA_Y12,
                          SA_EXP2F_C0, SA_EXP2F_C1, SA_EXP2F_C2, SA_EXP2F_SHIFT, -SA_EXP2F_SHIFT, 0.0};
    for (int i = 0; i < 12; i++) pk[i] = v[i];
}

int exclusive_scan_i64(int64_t *d_vals, int64_t n_plus_1, void *d_tmp, size_t tmp_bytes, size_t *tmp_needed, cudaStream_t st) {
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, d_vals, d_vals, (int)n_plus_1, st);
    if (tmp_needed) *tmp_needed = need;
    if (!d_tmp || tmp_bytes < need) return 1;
    return cub::DeviceScan::ExclusiveSum(d_tmp, need, d_vals, d_vals, (int)n_plus_1, st) == cudaSuccess ? 0 : 2;
}

// sa_align_rounds: the round records leave the device packed (a pair has 0..max_alns of them; at 10^7 pairs the dense
// [pair][max_alns] array would be gigabytes of pinned memory and PCIe traffic for mostly empty slots)
__global__ void rounds_count_kernel(const int32_t *nr, int64_t *off, long long m) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) off[i] = nr[i] > 0 ? nr[i] : 0;
    if (i == m) off[i] = 0;
}
__global__ void rounds_gather_kernel(const sa_round *rounds, const int32_t *nr, const int64_t *off, sa_round *packed, long long m, int max_alns) {
    // one thread per 16-byte half record: consecutive threads write consecutive halves
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = t / (2 * max_alns);
    const int h = (int)(t % (2 * max_alns));
    if (i >= m || (h >> 1) >= nr[i]) return;
    reinterpret_cast<uint4 *>(packed + off[i])[h] = reinterpret_cast<const uint4 *>(rounds + i * max_alns)[h];
}
void launch_rounds_count(const int32_t *nr, int64_t *off, long long m, cudaStream_t st) {
    rounds_count_kernel<<<(unsigned)((m + 1 + 255) / 256), 256, 0, st>>>(nr, off, m);
}
void launch_rounds_gather(const sa_round *rounds, const int32_t *nr, const int64_t *off, sa_round *packed, long long m, int max_alns, cudaStream_t st) {
    const long long threads = m * 2 * max_alns;
    if (threads > 0) rounds_gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(rounds, nr, off, packed, m, max_alns);
}

void launch_traceback(const TraceParams &p, cudaStream_t st) {
    const int threads = 64;
    const long long blocks = (p.n_problems + threads - 1) / threads;
    if (blocks > 0) trace_kernel<<<(unsigned)blocks, threads, 0, st>>>(p);
}

void launch_detect(const DetectParams &p, cudaStream_t st) {
    if (p.n_problems > 0) detect_kernel<<<(unsigned)p.n_problems, 256, 0, st>>>(p);
}
void launch_powf12(const float *x, long long n, float *out, int fast, cudaStream_t st) {
    cudaFuncSetAttribute(powf12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAB_SMEM_BYTES);
    powf12_kernel<<<148 * 4, 256, TAB_SMEM_BYTES, st>>>(x, n, out, fast);
}
void launch_powf12_range(uint32_t lo, uint32_t hi, float *out_fast, float *out_generic, cudaStream_t st) {
    cudaFuncSetAttribute(powf12_range_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAB_SMEM_BYTES);
    powf12_range_kernel<<<148 * 4, 256, TAB_SMEM_BYTES, st>>>(lo, hi, out_fast, out_generic);
}
void launch_prune_bound(uint32_t lo, uint32_t hi, unsigned long long *n_viol, float *min_ratio, cudaStream_t st) {
    cudaFuncSetAttribute(prune_bound_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAB_SMEM_BYTES);
    prune_bound_kernel<<<148 * 4, 256, TAB_SMEM_BYTES, st>>>(lo, hi, n_viol, min_ratio);
}
void launch_dfma_peak(double *sink, int iters, int n_ctas, cudaStream_t st) {
    dfma_peak_kernel<<<n_ctas, 256, 0, st>>>(sink, iters);
}
void launch_mufu_peak(float *sink, int iters, int n_ctas, cudaStream_t st) {
    mufu_peak_kernel<<<n_ctas, 256, 0, st>>>(sink, iters);
}
int max_clusters() {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CLUSTER_CTAS * 64);
    cfg.blockDim = dim3(CLUSTER_WARPS * 32);
    cfg.dynamicSmemBytes = fill_dyn_smem(CLUSTER_WARPS);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER_CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaFuncSetAttribute(fill_kernel<0, false, CLUSTER_WARPS, 1, CLUSTER_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)fill_dyn_smem(CLUSTER_WARPS));
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fill_kernel<0, false, CLUSTER_WARPS, 1, CLUSTER_CTAS>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int fill_ctas_per_sm() {
    int n = 0;
    cudaFuncSetAttribute(fill_kernel<0, false, 1, FILL_MINB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)fill_dyn_smem(FILL_WARPS));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fill_kernel<0, false, 1, FILL_MINB, 1>, FILL_WARPS * 32,
                                                  fill_dyn_smem(FILL_WARPS));
    return n > 0 ? n : 1;
}

}  // namespace sa
