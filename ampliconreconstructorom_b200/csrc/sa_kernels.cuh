// Device-side data structures shared by the kernels (sa_kernels.cu) and the C-ABI host layer (sa_capi.cu).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/segaligner_b200.h"

namespace sa {

constexpr int LB = 6;  // banded look-back of the reference (main.cpp:27); the -nl path uses fill_unbanded

// Per-label operand table, built once per map set on the host in exact fp32 with the reference's operation
// order (build_label_ops in sa_capi.cu):
//   d[k-1] = pos[l] - pos[l-k]                           interval terms of score_f (SegAligner.cpp:87)
//   f[k-1] = 5000f * (((0+prob[l-k+1])+...)+prob[l-1])   fp_t of score_f (:86,:90) for a look-back of k
// Entries whose predecessor l-k does not exist hold 0 (never selected: the predecessor score is -inf).
struct __align__(16) LabelOps {
    float d[LB];
    float f[LB];
};
static_assert(sizeof(LabelOps) == 48, "LabelOps must be 48 bytes");

struct MapSetDev {
    const float *pos;        // concatenated positions, n+1 per map
    const float *prob;       // concatenated non-collapse probabilities, n per map
    const LabelOps *ops;     // one per label (concatenated, n per map)
    const int64_t *pos_off;  // n_maps+1 offsets into pos; labels of map m start at pos_off[m]-m in ops/prob
    int32_t n_maps;
    int32_t max_labels;
};

// Wavefront units (see "coupled units" in sa_kernels.cu): a large problem is swept by U warps anywhere on the GPU; unit u owns
// the row strips u, u+U, u+2U, ... and hands the bottom rows of each to unit (u+1) % U through a tagged mailbox in L2.
struct CplUnit {
    int32_t problem;  // index into FillParams::problems
    int32_t u, U;
    int32_t slot0;    // mailbox / result slot of the problem's unit 0 (its units occupy slot0 .. slot0 + U - 1)
};
struct CplRes {  // a unit's backtrack-start candidate
    float score;
    int32_t j, q, pad;
    long long ord;
};

struct FillParams {
    MapSetDev contigs, segs;
    const sa_problem *problems;
    const int32_t *order;     // processing order (largest first) or nullptr
    const int64_t *cell_off;  // per problem offset (in cells) into S/code, STORE variants only
    long long n_problems;
    const uint32_t *used_bits;
    long long used_words_per_slot;
    float *S;       // stored score matrices
    uint8_t *code;  // stored predecessor codes: 0 = none, else 1 + (dj-1)*6 + (dq-1)
    sa_result *results;
    float *halo;    // per work-unit (warp, or CTA in the cooperative variant) scratch: halo_stride floats each
    long long halo_stride;
    unsigned long long *counter;  // work-queue head
    unsigned long long *stats;    // [0] fp64 replays executed, [1] cells rescanned exactly (pruning kernel), or nullptr
    float *cluster_mail;          // cluster variant: CLUSTER_MAIL_FLOATS floats per cluster
    int init_mode, start_mode, fitting;
    float *rowmax_s;  // detection only: per-row best positive score / its column (ties -> larger q)
    int32_t *rowmax_q;
    const int64_t *row_off;  // per problem offset into the rowmax arrays
    double pk[12];  // fp64 constants of the powf replay (kernel-parameter bank => direct DFMA operands)
    // ---- coupled units (CPL variant): n_problems counts UNITS, `order` is unused
    const CplUnit *units;
    unsigned long long *mbox;  // [slots][mbox_stride] entries {fp32 value, tag}: 6 per column, tag = producing strip * nx + q + 1
    long long mbox_stride;
    CplRes *cpl_res;           // [slots]
    int32_t *cpl_done;         // [slots] (entry slot0 of a problem counts its finished units)
    int32_t *cpl_error;        // set when a handshake timed out (never, unless the kernel was mis-launched)
    uint32_t cpl_epoch;        // launch number in the top 4 bits of every tag (the low 28 bits count the problem's steps)
    uint32_t cpl_spin_limit;   // polls after which a hand-over counts as lost (host: ~10 x the duration of the longest unit)
    // ---- STORE_MODE 3 ("rounds", sa_align_rounds): the whole `while` loop of run_SA_aln (main.cpp:268-300) per pair on one
    // warp.  Per-worker scratch (one slot per resident warp) that persists across the rounds of a pair:
    uint8_t *r_code;        // [workers][r_cells]         predecessor codes of the pair's whole matrix
    float *r_halo;          // [workers][r_halo_floats]   bottom 6 rows of EVERY strip: [strip][q][8]
    float *r_col2;          // [workers][2 * r_max_nb]    S[j][nx-2], S[j][nx-1] for every row
    float *r_lastrow;       // [workers][r_max_nx]        S[nb-1][q]
    uint32_t *r_used;       // [workers][r_used_words]    used contig labels of the pair (seed + accepted paths)
    int32_t *r_path;        // [workers][2 * r_max_path]  traceback of the current round: j's then q's
    float *r_ps;            // [workers][r_max_path]      scores along it
    long long r_cells, r_halo_floats;
    int r_max_nb, r_max_nx, r_used_words, r_max_path, r_max_alns;
    const float *r_thr;     // per problem: the segment's score threshold (curr_best_score / exp_thresh start value)
    sa_round *r_rounds;     // [n_problems][r_max_alns]
    int32_t *r_n_rounds;    // [n_problems]
    unsigned long long *r_arena_top;  // bump allocator over the packed path arena
    long long r_arena_cap;
    int32_t *r_pj, *r_pq;   // packed paths of all rounds of all pairs
    float *r_psc;
    int32_t *r_overflow;    // set when the arena was too small (host-sized for the worst case, so never)
    // sa_align_rounds_filtered: only tracebacks that can still become a file leave the device (r_flt_on)
    int r_flt_on, r_flt_tip, r_flt_min_len, r_flt_tipwin;
    float r_flt_mean, r_flt_mean_short, r_flt_stop_below;  // stop_below: -inf = never
    float r_win_gain, r_win_min, r_win_slope;  // tip end windows: per-step gain of the window sweeps, and the value a writable path reaches
};

struct TraceParams {
    MapSetDev contigs, segs;
    const sa_problem *problems;
    const int64_t *cell_off;
    long long n_problems;
    const float *S;          // stored scores, or nullptr: banded fill stored only the codes, scores are recomputed
    const uint8_t *code;     // banded: predecessor codes
    const int32_t *prev32;   // unbanded (-nl): packed predecessors i<<16|p, -1 = none
    const int64_t *path_off;
    sa_result *results;
    int32_t *path_j, *path_q;
    float *path_s;
    int32_t init_mode, swap;  // needed when S == nullptr (scores along the path are recomputed)
    int32_t count_only;  // 1: only measure the paths (results[k].path_len, len_out[k]); nothing is written
    int64_t *len_out;    // optional: path length per problem as int64 (input of the prefix sum)
};

struct DetectParams {
    MapSetDev contigs, segs;
    const sa_problem *problems;
    const int64_t *cell_off, *row_off, *scores_off;
    long long n_problems;
    float *S;
    const uint8_t *code;
    const int32_t *prev32;
    float *rowmax_s;
    int32_t *rowmax_q;
    const int32_t *n_want;
    float *scores_out;
    int32_t *n_got;
};

void launch_detect(const DetectParams &p, cudaStream_t st);
// store: 0 = scoring (nothing per cell), 1 = predecessor codes, 2 = codes + scores (+ per-row maxima when requested)
// shape: 0 = bulk (one warp per problem), 1 = cooperative CTA, 2 = cooperative cluster, 3 = coupled units (one warp per unit)
void launch_fill(const FillParams &p, int store, bool swap, int shape, int n_ctas, cudaStream_t st);
int max_clusters();
void launch_fill_unbanded(const FillParams &p, int32_t *prev32, bool swap, int n_ctas, cudaStream_t st);
void launch_traceback(const TraceParams &p, cudaStream_t st);
void launch_rounds_count(const int32_t *nr, int64_t *off, long long m, cudaStream_t st);
void launch_rounds_gather(const sa_round *rounds, const int32_t *nr, const int64_t *off, sa_round *packed, long long m, int max_alns, cudaStream_t st);
// in-place exclusive prefix sum of n+1 int64 values (the last input value is ignored: out[n] = total)
int exclusive_scan_i64(int64_t *d_vals, int64_t n_plus_1, void *d_tmp, size_t tmp_bytes, size_t *tmp_needed, cudaStream_t st);
void launch_powf12(const float *x, long long n, float *out, int fast, cudaStream_t st);
void launch_powf12_range(uint32_t lo, uint32_t hi, float *out_fast, float *out_generic, cudaStream_t st);
void launch_prune_bound(uint32_t lo, uint32_t hi, unsigned long long *n_viol, float *min_ratio, cudaStream_t st);
void launch_mufu_peak(float *sink, int iters, int n_ctas, cudaStream_t st);
void launch_dfma_peak(double *sink, int iters, int n_ctas, cudaStream_t st);
int fill_ctas_per_sm();
void fill_powf_constants(double *pk);
// Layout of the powf tables in shared memory (see sa_kernels.cu): 3 = lane-replicated, bank-conflict free, 68 KB per
// CTA => one 12-warp CTA per SM; 1 = single copy, 16 KB, three 4-warp CTAs per SM.  Either way 12 warps x 168
// registers per SM (168 lets ptxas interleave the 6 powf chains of a look-back row; measured best on B200).
// SA_PRUNE 1 (default): exact pruning -- fp32 upper bounds for all 36 candidates of a cell, ONE bit-exact fp64 replay
// for the winner, whole-cell exact fallback on (near-)ties.  0: the dense kernel (36 replays per cell).
#ifndef SA_PRUNE
#define SA_PRUNE 1
#endif
#ifndef SA_TAB2D
#define SA_TAB2D (SA_PRUNE ? 1 : 3)
#endif
#ifndef SA_FILL_MINB
#define SA_FILL_MINB (SA_PRUNE ? 3 : (SA_TAB2D == 3 ? 1 : 3))
#endif
#ifndef SA_FILL_WARPS
#define SA_FILL_WARPS (SA_PRUNE ? 8 : (SA_TAB2D == 3 ? 12 : 4))
#endif
constexpr int FILL_MINB = SA_FILL_MINB;   // resident CTAs per SM the fill kernel is compiled for
constexpr int CLUSTER_CTAS = 8;   // CTAs (SMs) per cluster of the cluster-cooperative variant used for huge problems
constexpr int CLUSTER_MAIL_FLOATS = 16 + CLUSTER_CTAS * 16 + CLUSTER_CTAS * 8;  // per-cluster L2 mailbox
#ifndef SA_COOP_WARPS
#define SA_COOP_WARPS (SA_PRUNE ? 8 : 16)
#endif
#ifndef SA_COOP_MINB
#define SA_COOP_MINB (SA_PRUNE ? 3 : 1)
#endif
constexpr int COOP_WARPS = SA_COOP_WARPS;  // warps of the cooperative (one problem per CTA) variant used for large problems
constexpr int COOP_MINB = SA_COOP_MINB;    // resident cooperative CTAs per SM (register budget 65536 / (MINB * 32 * warps))
constexpr int CLUSTER_WARPS = 16;  // warps per CTA of the cluster shape: its step barrier is a cluster barrier (~2 us), so
                                   // huge problems want the widest wavefront (8 SMs x 16 warps), one CTA per SM
constexpr int FILL_WARPS = SA_FILL_WARPS;  // warps per CTA of the fill kernel (one problem per warp at a time)

}  // namespace sa
