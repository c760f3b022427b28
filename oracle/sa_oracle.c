/* TEST INFRASTRUCTURE ONLY -- see sa_oracle.h.  Plain-C restatement of the SegAligner DP hot path.
 * Each function cites the reference lines (relative to /root/reference/SegAligner/) it follows.
 * Arithmetic is fp32 with one rounding per operation (compiled -ffp-contract=off, no -march, like the
 * reference Makefile:2-3), and pow is the host libm powf -- the same function the reference calls. */
#include "sa_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

float sao_powf12(float x) { return powf(x, 1.2f); }

/* Exhaustive check helper: for every float bit pattern in [lo, hi) compare the host libm powf(x, 1.2f) the reference
 * links (SegAligner.cpp:88) with two device result arrays (got_a/got_b, hi-lo entries each), bit for bit (NaN matches
 * NaN).  Returns the number of mismatching patterns; *first_bad receives the first one (or 0xffffffff). */
long long sao_powf12_cmp_range(unsigned lo, unsigned hi, const float *got_a, const float *got_b, unsigned *first_bad) {
    long long bad = 0;
    if (first_bad) *first_bad = 0xffffffffu;
    for (unsigned long long i = 0; i < (unsigned long long)hi - lo; i++) {
        unsigned bits = lo + (unsigned)i, e, a, b;
        float x, r;
        memcpy(&x, &bits, 4);
        r = powf(x, 1.2f);
        memcpy(&e, &r, 4);
        memcpy(&a, &got_a[i], 4);
        memcpy(&b, &got_b[i], 4);
        if (r != r) { /* NaN: any NaN on the device side matches */
            if (got_a[i] == got_a[i] || got_b[i] == got_b[i]) { if (!bad && first_bad) *first_bad = bits; bad++; }
        } else if (a != e || b != e) {
            if (!bad && first_bad) *first_bad = bits;
            bad++;
        }
    }
    return bad;
}

/* SegAligner.cpp:40-68.  left_0 = 1; nc_k = f(pos[k]-pos[k-1]); out[k-1] = left_{k-1}*nc_k; the last label
 * gets left_{last}*nc_last (:65) where nc_last is whatever `non_collapse` holds -- for a 1-label map that is
 * the value carried over from the previous map (:43). */
void sao_non_collapse_probs(const float *pos, int n1, int gen, float *carry, float *out) {
    int y_len = n1 - 1;
    float non_collapse = *carry;
    float left_prev = 1.0f; /* get<0>(pairs[i-1]) */
    int o = 0;
    for (int i = 1; i < y_len; ++i) {
        float dist = pos[i] - pos[i - 1];
        if (gen == 1) {
            non_collapse = fminf(1.0f, powf(dist, 4) / powf(1950, 4)); /* :53 */
        } else {
            if (dist < 475) {
                non_collapse = 0.0f; /* :56 */
            } else {
                non_collapse = (float)fmin(1.0f, (dist - 450) / 1500.0); /* :58, double promotion */
            }
        }
        out[o++] = left_prev * non_collapse; /* :63 */
        left_prev = non_collapse;            /* :61 */
    }
    if (y_len >= 1) out[o++] = left_prev * non_collapse; /* :65 */
    *carry = non_collapse;
}

/* SAHelper.cpp:72-78 */
void sao_reverse_map(const float *pos, int n1, float *out) {
    int o = 0;
    float back = pos[n1 - 1];
    for (int j = n1 - 2; j > -1; --j) out[o++] = back - pos[j];
    out[o] = back;
}

/* SegAligner.cpp:83-92 */
float sao_score_f(const float *b, const float *x, int i, int j, int p, int q, const float *xprobs) {
    float exp_x_labels = 0.0f;
    for (int k = p + 1; k < q; k++) exp_x_labels = exp_x_labels + xprobs[k]; /* std::accumulate, :86 */
    float discrep = fabsf((b[j] - b[i]) - (x[q] - x[p]));
    float delta = powf(discrep, 1.2f);
    float fn_t = 5000.0f * (float)(j - (i + 1));
    float fp_t = 5000.0f * exp_x_labels;
    return 10000.0f - (fn_t + fp_t + delta);
}

/* SegAligner.cpp:254-287.  The reference's vectors are value-initialised to 0 (main.cpp:137,190,269). */
void sao_init(float *S, int nb, int nx, int init_mode) {
    for (size_t k = 0; k < (size_t)nb * nx; k++) S[k] = 0.0f;
    if (init_mode == SAO_INIT_SEMIGLOBAL) { /* :254-269 */
        for (int j = 1; j < nb; j++)
            for (int q = 2; q < nx; q++) S[(size_t)j * nx + q] = -INFINITY;
    } else if (init_mode == SAO_INIT_FITTING) { /* :280-287 */
        for (int j = 1; j < nb; j++)
            for (int q = 1; q < nx; q++) S[(size_t)j * nx + q] = -INFINITY;
    }
}

/* SegAligner.cpp:95-133 */
void sao_dp_align(float *S, int *prev, const float *b, int nb, const float *x, int nx, int fitting,
                  int lookback, const unsigned char *used, const float *xprobs, const float *bprobs, int swap_b_x) {
    for (int j = 0; j < nb; j++) {
        int i_start = j - lookback > 0 ? j - lookback : 0;
        for (int q = 0; q < nx; ++q) {
            size_t c = (size_t)j * nx + q;
            prev[2 * c] = -1;
            prev[2 * c + 1] = -1;
            if (used && used[j]) continue; /* :104 */
            int p_start = q - lookback > 0 ? q - lookback : 0;
            for (int i = i_start; i < j; i++) {
                for (int p = p_start; p < q; p++) {
                    if (fitting && i == 0 && p > 0) break; /* :111 */
                    float newScore;
                    if (!swap_b_x) newScore = S[(size_t)i * nx + p] + sao_score_f(b, x, i, j, p, q, xprobs);
                    else newScore = S[(size_t)i * nx + p] + sao_score_f(x, b, p, q, i, j, bprobs);
                    if (newScore > S[c]) { /* strict, :123 */
                        S[c] = newScore;
                        prev[2 * c] = i;
                        prev[2 * c + 1] = p;
                    }
                }
            }
        }
    }
}

/* SegAligner.cpp:172-195.  x_len-3 = nx-2 (x_len = nx+1). */
void sao_start_sg(const float *S, int nb, int nx, int *jq) {
    float best = -INFINITY;
    jq[0] = -1;
    jq[1] = -1;
    int j_end = nb - 1;
    for (int q = 0; q < nx; q++)
        if (S[(size_t)j_end * nx + q] > best) {
            best = S[(size_t)j_end * nx + q];
            jq[0] = j_end;
            jq[1] = q;
        }
    for (int j = 0; j < nb; j++)
        for (int q = nx - 2; q < nx; q++) {
            if (q < 0) continue; /* reference reads S[j][-1] here when nx==1 (UB); defined as "skip" */
            if (S[(size_t)j * nx + q] > best) {
                best = S[(size_t)j * nx + q];
                jq[0] = j;
                jq[1] = q;
            }
        }
}

/* SegAligner.cpp:142-154 */
void sao_start_local(const float *S, int nb, int nx, int *jq) {
    float best = -INFINITY;
    jq[0] = -1;
    jq[1] = -1;
    for (int j = 0; j < nb; j++)
        for (int q = 0; q < nx; q++)
            if (S[(size_t)j * nx + q] > best) {
                best = S[(size_t)j * nx + q];
                jq[0] = j;
                jq[1] = q;
            }
}

/* SAHelper.cpp:83-95 */
int sao_traceback(const float *S, const int *prev, int nx, int j, int q, int *path_jq, float *path_s) {
    int n = 0;
    while (j != -1) {
        size_t c = (size_t)j * nx + q;
        path_jq[2 * n] = j;
        path_jq[2 * n + 1] = q;
        path_s[n] = S[c];
        n++;
        int nj = prev[2 * c], nq = prev[2 * c + 1];
        j = nj;
        q = nq;
    }
    return n;
}

typedef struct {
    float s;
    int j, q;
} sao_cell;
static int cell_cmp(const void *a, const void *b) { /* tuple<float,int,int> operator< , SegAligner.cpp:219 */
    const sao_cell *u = (const sao_cell *)a, *v = (const sao_cell *)b;
    if (u->s < v->s) return -1;
    if (u->s > v->s) return 1;
    if (u->j != v->j) return u->j < v->j ? -1 : 1;
    if (u->q != v->q) return u->q < v->q ? -1 : 1;
    return 0;
}

/* SegAligner.cpp:198-233.  The index is decremented before use (:224) so the single best cell is always
 * skipped.  The reference reads all_scores[-1] when it runs out of candidates (UB); here the walk stops. */
int sao_multi_backtrack(float *S, const int *prev, int nb, int nx, int n, float *scores_out) {
    size_t cap = 0;
    for (size_t k = 0; k < (size_t)nb * nx; k++) cap += S[k] > 0;
    sao_cell *all = (sao_cell *)malloc((cap ? cap : 1) * sizeof(sao_cell));
    size_t m = 0;
    for (int j = 0; j < nb; j++)
        for (int q = 0; q < nx; q++)
            if (S[(size_t)j * nx + q] > 0) {
                all[m].s = S[(size_t)j * nx + q];
                all[m].j = j;
                all[m].q = q;
                m++;
            }
    qsort(all, m, sizeof(sao_cell), cell_cmp);
    long last = (long)m - 1;
    int got = 0;
    for (int it = 0; it < n; it++) {
        int hit = 0;
        while (!hit && last > -1) {
            last -= 1;
            if (last < 0) break; /* defined stop (reference: UB) */
            int j = all[last].j, q = all[last].q;
            float cur = S[(size_t)j * nx + q];
            if (cur > 0) {
                hit = 1;
                /* zero_out_aln :198-205 */
                while (j != -1) {
                    size_t c = (size_t)j * nx + q;
                    S[c] = -INFINITY;
                    int nj = prev[2 * c], nq = prev[2 * c + 1];
                    j = nj;
                    q = nq;
                }
                scores_out[got++] = cur;
            }
        }
    }
    free(all);
    return got;
}

float sao_score_problem(const float *b, int nb, const float *x, int nx, const float *xprobs, const float *bprobs,
                        int lookback, int swap_b_x, int start_mode, int *jq) {
    float *S = (float *)malloc(sizeof(float) * (size_t)nb * nx);
    int *prev = (int *)malloc(sizeof(int) * 2 * (size_t)nb * nx);
    sao_init(S, nb, nx, SAO_INIT_SEMIGLOBAL); /* main.cpp:141 "handles both cases" */
    sao_dp_align(S, prev, b, nb, x, nx, 0, lookback, NULL, xprobs, bprobs, swap_b_x);
    int st[2];
    if (start_mode == 0) sao_start_sg(S, nb, nx, st);
    else sao_start_local(S, nb, nx, st);
    float sc = (st[0] >= 0) ? S[(size_t)st[0] * nx + st[1]] : -INFINITY;
    if (jq) {
        jq[0] = st[0];
        jq[1] = st[1];
    }
    free(S);
    free(prev);
    return sc;
}

void sao_score_batch(int64_t n, const int *pair_contig, const int *pair_seg, const int64_t *c_off, const float *c_pos,
                     const float *c_prob, const int64_t *s_off, const float *s_pos, const float *s_prob,
                     int lookback, int swap_b_x, int start_mode, float *scores_out) {
    for (int64_t k = 0; k < n; k++) {
        int c = pair_contig[k], s = pair_seg[k];
        int nb = (int)(c_off[c + 1] - c_off[c]) - 1, nx = (int)(s_off[s + 1] - s_off[s]) - 1;
        scores_out[k] = sao_score_problem(c_pos + c_off[c], nb, s_pos + s_off[s], nx, s_prob + (s_off[s] - s),
                                          c_prob + (c_off[c] - c), lookback, swap_b_x, start_mode, NULL);
    }
}

static int float_cmp(const void *a, const void *b) {
    float u = *(const float *)a, v = *(const float *)b;
    return u < v ? -1 : (u > v ? 1 : 0);
}

/* main.cpp:43-105 for one segment.  m = total contig labels (:69-72), n = this segment's labels (:76). */
float sao_score_threshold(const float *scores, int64_t n_sc, int n_contigs, int n_detect_scores, int m, int n,
                          double pvalue) {
    float *s = (float *)malloc(sizeof(float) * (size_t)(n_sc ? n_sc : 1));
    memcpy(s, scores, sizeof(float) * (size_t)n_sc);
    qsort(s, (size_t)n_sc, sizeof(float), float_cmp); /* :59 */
    int right_cutoff = 25;
    int left_cut = (int)round(0.15 * fmin((double)n_contigs, (double)n_detect_scores)); /* :66 */
    double E_cutoff = -log(1 - pvalue);
    int64_t first = n_sc - (right_cutoff + left_cut); /* :77 */
    int64_t last = n_sc - 1 - right_cutoff;          /* :79 */
    if (n < 12) last -= 25;                           /* :80 */
    if (first < 0 || last > n_sc || last < first) {   /* reference: out-of-range iterators (UB); observed "-nan" */
        free(s);
        return -NAN;
    }
    int64_t cnt = last - first;
    /* e_vals = cnt..1, then int-truncated natural log (:85-90) */
    double nn = (double)cnt;
    double sx = 0.0, sy = 0.0;
    for (int64_t k = 0; k < cnt; k++) sx += s[first + k];
    for (int64_t k = 0; k < cnt; k++) sy += (int)log((double)(cnt - k));
    double xbar = sx / nn, ybar = sy / nn; /* SegAligner.cpp:304-319 */
    double cov = 0.0, var = 0.0;
    for (int64_t k = 0; k < cnt; k++) {
        double xv = s[first + k], yv = (int)log((double)(cnt - k));
        cov += (xv - xbar) * (yv - ybar);
        var += (xv - xbar) * (xv - xbar);
    }
    double slope = cov / var;
    double intercept = ybar - slope * xbar;
    int mn = (int)((unsigned)m * (unsigned)n); /* int product, wraps like the reference's (:97) */
    double K = exp(intercept - log((double)mn));
    double S_cutoff = -log(E_cutoff / (K * m * n)) / (-slope); /* :100 */
    free(s);
    return (float)S_cutoff;
}
