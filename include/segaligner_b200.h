/* segaligner_b200.h -- C ABI of the B200-native SegAligner DP engine (libsegaligner_b200.so).
 *
 * The reference has no in-process FFI for this path: its boundary is the SegAligner executable
 * (/root/reference/SegAligner/main.cpp:453) that ARAlignDetect.py:283-287, OMPathFinder.py:416-420 and
 * align_to_candidate_path.py:17-22 exec.  The drop-in is therefore the `SegAligner` executable built from
 * ampliconreconstructorom_b200/csrc/host/, and THIS header is the thin C layer between that C++ host and the
 * sm_100a kernels.  Each entry point below names the reference function(s) whose work it replaces, batched
 * over many independent (contig x segment) problems.  Plain pointers and sizes only; no exceptions cross
 * the boundary; every function returns 0 on success or a negative SA_E* code (sa_last_error() has text).
 *
 * Conventions (identical to the reference): a map with n labels is n+1 floats [pos_0..pos_{n-1}, length]
 * (SAHelper.cpp:44).  b = contig map = DP rows j; x = segment map = DP columns q.  All scoring arithmetic is
 * fp32 in the reference's operation order, powf(discrep,1.2f) is reproduced bit-for-bit (fp64 inside).
 */
#ifndef SEGALIGNER_B200_H
#define SEGALIGNER_B200_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SA_OK 0
#define SA_EINVAL -1  /* bad argument */
#define SA_ECUDA -2   /* CUDA runtime error (text in sa_last_error) */
#define SA_ENOMEM -3  /* device or host allocation failed */
#define SA_ENODEV -4  /* no usable sm_100 device */

/* S-matrix initialisation: init_semiglobal_aln / init_local_aln / init_fitting_aln (SegAligner.cpp:254-287) */
enum { SA_INIT_SEMIGLOBAL = 0, SA_INIT_LOCAL = 1, SA_INIT_FITTING = 2 };
/* traceback start: sg_aln_backtrack_start (SegAligner.cpp:172), local_backtrack_start (:142),
 * fixed (last contig label, last segment label) as run_SA_fitting does (main.cpp:197-199) */
enum { SA_START_SEMIGLOBAL = 0, SA_START_LOCAL = 1, SA_START_END = 2 };
enum { SA_SIDE_CONTIG = 0, SA_SIDE_SEG = 1 };

typedef struct sa_ctx sa_ctx; /* one per GPU; not thread-safe: use one host thread per context */

/* Arguments of dp_align (SegAligner.cpp:95-97) that are common to a batch. */
typedef struct sa_mode {
    int32_t init_mode;  /* SA_INIT_* */
    int32_t start_mode; /* SA_START_* */
    int32_t fitting;    /* dp_align's fitting_aln flag (SegAligner.cpp:111) */
    int32_t lookback;   /* 6 = banded kernels; anything larger = the -nl path (main.cpp:475-478) */
    int32_t swap_b_x;   /* dp_align's swap_b_x (SegAligner.cpp:115-121); set by -detection */
} sa_mode;

/* One DP problem: indices into the uploaded map sets. used_slot selects a used-label bitmap (or -1). */
typedef struct sa_problem {
    int32_t contig;
    int32_t seg;
    int32_t used_slot;
    int32_t reserved;
} sa_problem;

/* Backtrack start and its score (S[j][q]); j = q = -1 when no cell qualifies (as the reference returns). */
typedef struct sa_result {
    float score;
    int32_t j;
    int32_t q;
    int32_t path_len; /* filled by sa_align_batch */
} sa_result;

int sa_device_count(void);
int sa_ctx_create(int device, sa_ctx **out);
void sa_ctx_destroy(sa_ctx *ctx);
const char *sa_last_error(const sa_ctx *ctx);

/* Upload one side's maps.  off[n_maps+1] indexes pos (n_i+1 floats per map); prob holds the non-collapse
 * probabilities (non_collapse_probs, SegAligner.cpp:40-68), n_i floats per map, concatenated in map order
 * (map m's probabilities start at off[m]-m).  Host buffers; copied before return. */
int sa_set_maps(sa_ctx *ctx, int side, int32_t n_maps, const int64_t *off, const float *pos, const float *prob);

/* Phase-1 scoring of a batch: for each problem init S, dp_align, pick the backtrack start, return its score.
 * Replaces the body of run_SA_score's double loop (main.cpp:124-163, non-detection branch).
 * Host buffers in, host buffers out; H2D/D2H copies are inside the call. */
int sa_score_batch(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n, sa_result *results);

/* The same work split into its stages, so that a caller (or a benchmark) can keep inputs resident in HBM:
 *   sa_plan_create   validates + uploads a problem list and its processing order (H2D; synchronises).  Large
 *                    problems are routed to the cooperative kernel, the rest largest-first to per-warp workers.
 *   sa_plan_score    launches the scoring kernels on the context's stream; no host<->device traffic; async.
 *   sa_plan_results  copies the sa_result array back (D2H; synchronises).
 * sa_score_batch == create + score + results + destroy. */
typedef struct sa_plan sa_plan;
int sa_plan_create(sa_ctx *ctx, const sa_problem *problems, int64_t n, sa_plan **out);
void sa_plan_destroy(sa_ctx *ctx, sa_plan *plan);
int sa_plan_score(sa_ctx *ctx, sa_plan *plan, const sa_mode *mode);
int sa_plan_results(sa_ctx *ctx, sa_plan *plan, sa_result *results);
int sa_plan_info(const sa_plan *plan, int64_t *n_bulk, int64_t *n_large, double *total_cells);
int sa_sync(sa_ctx *ctx);

/* Alignment of a batch (one round of run_SA_aln's while loop, main.cpp:268-291, or run_SA_fitting,
 * main.cpp:190-202): init S, dp_align honouring each problem's used contig labels, backtrack start,
 * get_aln_list traceback (SAHelper.cpp:83-95).
 *   used_bits / used_words_per_slot: bitmap pool; problem k with used_slot s>=0 skips contig label j when bit
 *   j of words [s*used_words_per_slot, (s+1)*used_words_per_slot) is set (dp_align's used_labels set).
 *   path_off[n+1]: caller-chosen offsets; problem k may write up to path_off[k+1]-path_off[k] entries
 *   (min(nb,nx) always suffices).  path_j/path_q/path_s receive the traceback end->start, like aln_list. */
int sa_align_batch(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n,
                   const uint32_t *used_bits, int64_t used_words_per_slot, int64_t n_used_slots,
                   const int64_t *path_off, sa_result *results, int32_t *path_j, int32_t *path_q, float *path_s);

/* The same, with the tracebacks PACKED: path k is path_j/q/s[start[k] .. start[k] + results[k].path_len), end -> start.
 * All pointers refer to pinned host memory owned by the context and stay valid until the next sa_align_batch* call on it;
 * only sum(path_len) entries cross PCIe (the capacity-offset form above moves min(nb,nx) per problem).  Used by the
 * executable's run_SA_aln rounds (main.cpp:213-358), where tens of millions of pairs are aligned per round. */
typedef struct sa_packed_paths {
    int64_t n, total;          /* problems, packed path entries */
    const sa_result *results;  /* n */
    const int64_t *start;      /* n + 1 */
    const int32_t *path_j, *path_q;
    const float *path_s;
} sa_packed_paths;
int sa_align_batch_packed(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n, const uint32_t *used_bits,
                          int64_t used_words_per_slot, int64_t n_used_slots, sa_packed_paths *out);

/* All alignment rounds of a batch of (segment, contig) pairs in ONE call: the whole per-pair loop of run_SA_aln
 * (main.cpp:262-300) -- `while (curr_best_score >= exp_thresh && count < max_alns)`: init S (semi-global), dp_align
 * honouring the pair's used contig labels, sg_aln_backtrack_start, get_aln_list, compute_partial_score_threshold
 * (SegAligner.cpp:23-36), and on `best > exp_thresh` count++ and the path's contig labels join the used set -- runs on the
 * device, one pair per warp.  Round 1 fills the whole matrix; a later round recomputes only the rows its predecessor's
 * path can influence ([first path label, last path label + 6*(nx-1)], exact by the look-back cone of dp_align) from
 * per-pair state kept in HBM (predecessor codes, strip-boundary rows, last row and last two columns).
 *   thr[k]        the segment's score threshold (start value of curr_best_score and exp_thresh, main.cpp:262,266)
 *   used_slot     seed bitmap of the pair (tip phase: labels already taken on the contig) or -1; the seed is not modified
 * Output (pinned host memory owned by the context, valid until the next sa_align* call): n_rounds[k] rounds per pair
 * (0 when thr is NaN), packed in pair order: pair k's records are rounds[round_off[k] .. round_off[k] + n_rounds[k]) (a
 * batch of 10^7 pairs, most of which end after a round or two, returns megabytes instead of gigabytes); round r's traceback
 * (end -> start, like aln_list) is
 * path_j/q/s[path_off .. path_off + path_len).  The last round of a pair is the one that failed `best > exp_thresh`
 * (or the max_alns-th accepted one).  Semi-global mode, lookback 6, no swap_b_x, every segment >= 2 labels; a pair whose
 * matrix exceeds the per-worker scratch gets n_rounds = -1 (the caller falls back to sa_align_batch_packed rounds).
 *
 * sa_align_rounds_filtered: the same, but a round's traceback is only RETURNED when the alignment could still be written
 * to a file -- the round was accepted and passes the filters of run_SA_aln that need nothing but the path's two ends, its
 * length and its total score (main.cpp:303-325: tip alignments must touch a contig end, minimum label count, mean score
 * per step; the mean is tested with 0.1 % slack, the caller applies the exact mean / median / minimum tests to what comes
 * back).  Every other round has path_off = -1: its record (score, start, path_len, j_first) is all the caller's replay of
 * the loop needs.  On C2-shaped runs this keeps > 99 % of the traceback entries on the device. */
typedef struct sa_round {
    float score;       /* S at the backtrack start = curr_best_score of the round */
    int32_t j, q;      /* backtrack start = LAST contig / segment label of the path */
    int32_t path_len;
    int64_t path_off;  /* -1: the traceback was not returned (sa_align_rounds_filtered) */
    int32_t j_first;   /* FIRST contig label of the path (what compute_partial_score_threshold needs besides j) */
    int32_t reserved;
} sa_round;
typedef struct sa_rounds_filter {
    int32_t tip;           /* 1: run_SA_aln's tip_aln test (main.cpp:303-308) */
    int32_t min_len;       /* fewer path entries cannot be written (main.cpp:310-312) */
    float mean_min;        /* score / (path_len - 1) below this cannot be written, paths of >= 5 entries (main.cpp:322) */
    float mean_min_short;  /* the same for paths of < 5 entries (main.cpp:318) */
    float stop_below;      /* > 0: a pair's loop ends after the first round whose best score is below this value.  Every later
                            * round's best score is <= it (masking labels only removes candidates; fp32 '+' and max are monotone),
                            * so when no alignment scoring less can be written, the rounds left out are unobservable.  The round
                            * that fell below is still reported; n_rounds counts up to it.  <= 0: run the loop to its own end. */
    int32_t tip_windows;   /* 1 (with tip = 1): before a pair's first full sweep only the strips that an alignment touching a
                            * contig end AND passing the filters can lie in are swept (rows 0 .. 2 + s, counting only paths that
                            * start in rows 0..2, and rows >= nb - 4 - s, s = (nx - 1)(1 + c / 5000): a skipped label costs 5000
                            * of a budget of c per step), with a per-step gain G = min(10000, 1.1 (min_len - 1) c) in place of
                            * score_f's 10000, c = 10000 - 0.999 min(mean_min, mean_min_short): a path that passes
                            * the length and mean-score filters is then worth >= (G - c)(min_len - 1), and masking labels only
                            * lowers cell values, so if no backtrack start of the windows reaches that value, no round of the
                            * pair can be written and the pair is reported with n_rounds = -2 without having been aligned.
                            * Pairs whose two windows overlap (short contigs) are aligned as usual. */
} sa_rounds_filter;
typedef struct sa_rounds_out {
    int64_t n;
    int32_t max_alns;
    int32_t reserved;
    int64_t total;             /* packed path entries */
    const int32_t *n_rounds;   /* n (-1: not aligned, too large for the scratch; -2: certified unwritable, see tip_windows) */
    const sa_round *rounds;    /* packed: sum over k of max(n_rounds[k], 0) records */
    const int32_t *path_j, *path_q;
    const float *path_s;
    const int64_t *round_off;  /* n + 1: where pair k's records start in rounds */
} sa_rounds_out;
int sa_align_rounds(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n, const float *thr, int32_t max_alns,
                    const uint32_t *used_bits, int64_t used_words_per_slot, int64_t n_used_slots, sa_rounds_out *out);
int sa_align_rounds_filtered(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n, const float *thr,
                             int32_t max_alns, const uint32_t *used_bits, int64_t used_words_per_slot, int64_t n_used_slots,
                             const sa_rounds_filter *filter, sa_rounds_out *out);
/* Optional: pins the host buffers that sa_align_rounds* calls of up to n_pairs pairs return their results in, ahead of time
 * (pinning costs 0.3 - 1 s per GB; growing a pinned buffer inside a call waits for the device).  The one call that may be
 * made from ANOTHER host thread while sa_score_batch / sa_detect_batch runs on the context -- run_SA_score is when the host
 * has nothing else to do (main.cpp:531-551) -- but not beside any sa_align* call. */
int sa_prepare_rounds(sa_ctx *ctx, int64_t n_pairs, int32_t max_alns);

/* Detection-mode scoring (run_SA_score with detection=true, main.cpp:155-159): semiglobal init, dp_align,
 * then multi_backtrack_scores (SegAligner.cpp:208-233) harvesting n_want[k] path scores per problem in the
 * reference's order (including its skip of the single best cell).  scores_off[n+1] gives each problem's slot
 * range in scores_out (capacity >= n_want[k]); n_got[k] receives the number actually produced. */
int sa_detect_batch(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n, const int32_t *n_want,
                    const int64_t *scores_off, float *scores_out, int32_t *n_got);

/* Full stored matrices of ONE problem (testing / debugging aid mirroring dp_align's S and previous outputs):
 * S is nb*nx floats row-major, prev is 2 ints per cell ({-1,-1} = none). */
int sa_fill_one(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problem, const uint32_t *used_bits,
                int64_t used_words, float *S, int32_t *prev, sa_result *result);

/* Bit-exact powf(x, 1.2f) on the device for n host floats (test hook for the score_f delta term). */
int sa_powf12(sa_ctx *ctx, const float *x, int64_t n, float *out, int use_fast_path);
/* Exhaustive device check: evaluates both device paths for every float bit pattern in [lo_bits, hi_bits) and
 * counts results that differ from `expected` (host libm powf supplied by the caller), chunk by chunk. */
int sa_powf12_range(sa_ctx *ctx, uint32_t lo_bits, uint32_t hi_bits, float *out_fast, float *out_generic);
/* Certification of the exact-pruning bound of the fill kernel: for every float bit pattern x in [lo_bits, hi_bits)
 * checks on the device that the fp32 lower bound of powf(x, 1.2f) used to prune candidates really is <= the bit-exact
 * replay (SegAligner.cpp:88).  *n_violations must come back 0; *min_ratio = smallest exact/bound seen (slack). */
int sa_check_prune_bound(sa_ctx *ctx, uint32_t lo_bits, uint32_t hi_bits, uint64_t *n_violations, float *min_ratio);

/* Last kernel timings of the most recent batch call on this context, in milliseconds (CUDA events). */
typedef struct sa_timing {
    float fill_ms;      /* DP fill kernel(s) */
    float post_ms;      /* traceback / multi-backtrack kernels */
    int32_t n_launches; /* kernels launched by the call */
    int32_t reserved;
} sa_timing;
int sa_get_timing(sa_ctx *ctx, sa_timing *t);
/* Work actually executed by the fill kernels of the last batch call: the fill prunes the 36 predecessor candidates
 * of a cell (SegAligner.cpp:108-127) with an fp32 bound and replays powf bit-exactly only for the winner.
 * n_replays = bit-exact powf replays (about one per cell), n_rescans = cells rescanned exactly after a (near-)tie. */
int sa_get_prune_stats(sa_ctx *ctx, uint64_t *n_replays, uint64_t *n_rescans);

/* Bytes this context has copied host->device and device->host since sa_ctx_create (every copy the library makes is
 * on the scoring / alignment / detection paths is counted: maps, label operand tables, problem lists, bitmaps, results,
 * packed tracebacks; the sa_fill_one and sa_digest helpers are not). */
int sa_get_traffic(sa_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes);

/* Microbenchmark: sustained fp64 FMA issue rate (DFMA instructions/s, whole GPU) used as roofline peak. */
int sa_measure_fp64_peak(sa_ctx *ctx, double *dfma_per_s);
/* Microbenchmark: sustained MUFU (lg2/ex2, the XU pipe) issue rate, thread-instructions/s over the whole GPU.  The
 * pruning fill kernel spends two MUFU per predecessor candidate (its fp32 score bound), which makes the XU pipe -- not
 * fp64 issue -- the binding resource of the executed work; bench.py reports the kernel against this peak. */
int sa_measure_mufu_peak(sa_ctx *ctx, double *mufu_per_s);

/* ---- host-side (CPU) preparation helpers: the executable's own code behind extern "C" -------------------------
 * None of these is on the DP hot path; they exist so that every caller prepares inputs identically.
 *   sa_host_non_collapse_probs  non_collapse_probs (SegAligner.cpp:40-68) over a map set in id order
 *   sa_host_make_reverse        make_reverse_cmap (SAHelper.cpp:61-81); two-call sizing: pass caps of 0 first
 *   sa_host_parse_cmap          parse_cmap (SAHelper.cpp:17-48); two-call sizing
 *   sa_host_score_thresholds    compute_score_thresholds (main.cpp:43-105), one float per segment (id order)
 *   sa_host_format_float        what `ostream << float` prints (the number format of every output file) */
int sa_host_non_collapse_probs(int32_t n_maps, const int64_t *off, const float *pos, int32_t inst_gen, float *prob_out);
int sa_host_make_reverse(int32_t n_maps, const int32_t *ids, const int64_t *off, const float *pos, int32_t min_map_len,
                         int32_t cap_maps, int64_t cap_pos, int32_t *out_ids, int64_t *out_off, float *out_pos,
                         int32_t *n_out, int64_t *n_pos_out);
int sa_host_parse_cmap(const char *path, int32_t cap_maps, int64_t cap_pos, int32_t *out_ids, int64_t *out_off,
                       float *out_pos, int32_t *n_out, int64_t *n_pos_out);
int sa_host_score_thresholds(int64_t n_rows, const int32_t *row_seg, const int32_t *row_contig, const float *row_score,
                             int32_t n_segs, const int32_t *seg_ids, const int64_t *seg_off, const float *seg_pos,
                             int32_t n_contigs, const int32_t *contig_ids, const int64_t *contig_off,
                             const float *contig_pos, double pvalue, int32_t n_detect_scores, float *thr_out);
int sa_host_format_float(float v, char *buf, int32_t cap);

/* ---- next row of the scope table: in-silico digestion, the producer of the CMAP inputs ------------
 * sa_digest replaces the motif scan of generate_cmap.py:119-123: finds every (overlapping) occurrence of `motif`
 * (4..8 letters ACGT, case-insensitive sequence) and of its reverse complement in seq[0..n), converts match starts
 * to label positions (start + offset on the forward strand, start + motif_len - offset on the reverse strand;
 * `offset` is the enzyme's cut offset + 1 as makeCMAP uses it) and returns them ascending and de-duplicated
 * (sorted(set(...))).  Positions can be < 0 or > n, exactly as in the reference.  Two-call sizing: with cap == 0 (or
 * too small) only *n_out is set, to an upper bound (raw match count).  Host buffers; H2D/D2H inside. */
int sa_digest(sa_ctx *ctx, const char *seq, int64_t n, const char *motif, int32_t motif_len, int32_t offset, int64_t cap,
              int64_t *pos_out, int64_t *n_out);

/* Device-time bracket on the context's stream (CUDA events): record slot 0 / 1, then read the elapsed ms. */
int sa_event_record(sa_ctx *ctx, int32_t slot);
int sa_event_elapsed_ms(sa_ctx *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif
