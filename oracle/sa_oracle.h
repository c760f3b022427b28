/* TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, fp32, host libm powf) of the SegAligner hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product (libsegaligner_b200.so / SegAligner) never links, imports or executes anything under oracle/.
 *
 * Parity status: PINNED -- every function here is checked in tests/test_oracle_vs_reference.py against
 * the unmodified reference compiled from /root/reference/SegAligner (oracle/_ref/libsa_ref.so) and
 * against the known-answer vectors of SURVEY.md section 8(c) stored under tests/golden/.
 *
 * Conventions: a map with n labels is n+1 floats [pos_0..pos_{n-1}, length] (SAHelper.cpp:44).
 * b = contig map (rows j, nb labels), x = segment map (columns q, nx labels).  S and prev are dense
 * row-major nb x nx (the reference allocates one more, never-touched row: main.cpp:137). */
#ifndef SA_ORACLE_H
#define SA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { SAO_INIT_SEMIGLOBAL = 0, SAO_INIT_LOCAL = 1, SAO_INIT_FITTING = 2 };

float sao_powf12(float x);
/* SegAligner.cpp:40-68, one map; *carry is the reference's `non_collapse` variable that survives across maps. */
void sao_non_collapse_probs(const float *pos, int n1, int gen, float *carry, float *out);
/* SAHelper.cpp:61-81, one map (n1 floats in, n1 floats out). */
void sao_reverse_map(const float *pos, int n1, float *out);
/* SegAligner.cpp:83-92 */
float sao_score_f(const float *b, const float *x, int i, int j, int p, int q, const float *xprobs);
/* SegAligner.cpp:254-287 */
void sao_init(float *S, int nb, int nx, int init_mode);
/* SegAligner.cpp:95-133; used = nb flags or NULL; prev = 2 ints per cell. S must be initialised. */
void sao_dp_align(float *S, int *prev, const float *b, int nb, const float *x, int nx, int fitting,
                  int lookback, const unsigned char *used, const float *xprobs, const float *bprobs, int swap_b_x);
/* SegAligner.cpp:172-195 / :142-154 */
void sao_start_sg(const float *S, int nb, int nx, int *jq);
void sao_start_local(const float *S, int nb, int nx, int *jq);
/* SAHelper.cpp:83-95; returns length; entries end->start */
int sao_traceback(const float *S, const int *prev, int nx, int j, int q, int *path_jq, float *path_s);
/* SegAligner.cpp:198-233; modifies S; returns count. Defined (non-UB) variant: stops when candidates run out. */
int sao_multi_backtrack(float *S, const int *prev, int nb, int nx, int n, float *scores_out);
/* whole phase-1 scoring problem: init semiglobal + dp_align + start (start_mode 0 sg, 1 local); returns score */
float sao_score_problem(const float *b, int nb, const float *x, int nx, const float *xprobs, const float *bprobs,
                        int lookback, int swap_b_x, int start_mode, int *jq);
/* Sum over a batch of problems given flattened map sets; used by bench.py's cpu_baseline leg.
 * pairs = (contig index, seg index). scores_out[n]. */
void sao_score_batch(int64_t n, const int *pair_contig, const int *pair_seg, const int64_t *c_off, const float *c_pos,
                     const float *c_prob, const int64_t *s_off, const float *s_pos, const float *s_prob,
                     int lookback, int swap_b_x, int start_mode, float *scores_out);
/* main.cpp:43-105 for one segment: scores (unsorted, n of them), returns threshold (NaN if window invalid). */
float sao_score_threshold(const float *scores, int64_t n, int n_contigs, int n_detect_scores, int m_total_labels,
                          int seg_labels, double pvalue);
/* host libm powf(x, 1.2f) vs two device result arrays over the float bit patterns [lo, hi); returns #mismatches */
long long sao_powf12_cmp_range(unsigned lo, unsigned hi, const float *got_a, const float *got_b, unsigned *first_bad);
#ifdef __cplusplus
}
#endif
#endif
