// TEST INFRASTRUCTURE ONLY (oracle/): extern "C" wrappers over the UNMODIFIED reference functions.
// This file contains no algorithm of its own: it marshals flat arrays into the std::map / std::vector
// containers the reference uses and calls the reference's own symbols, which are compiled from the
// sources where they lie under /root/reference/SegAligner (see oracle/Makefile).  The result,
// oracle/_ref/libsa_ref.so, is what pins oracle/sa_oracle.c ("parity pinned by the reference run here").
//
// Wrapped reference entry points (file:line in /root/reference/SegAligner):
//   score_f SegAligner.cpp:83, dp_align :95, non_collapse_probs :40, init_* :254-287,
//   sg_aln_backtrack_start :172, local_backtrack_start :142, multi_backtrack_scores :208,
//   compute_partial_score_threshold :23, compute_aln_median :321, compute_aln_min :345,
//   parse_cmap SAHelper.cpp:17, make_reverse_cmap :61, get_aln_list :83,
//   compute_score_thresholds main.cpp:43 (main.cpp is compiled with -Dmain=sa_ref_main).
#include <map>
#include <set>
#include <vector>
#include <array>
#include <tuple>
#include <string>
#include <cstring>
#include <cstdint>
#include <limits>
#include "SAHelper.h"
#include "SegAligner.h"

extern int n_detect_scores;  // main.cpp:24

namespace {
std::map<int, std::vector<float>> to_map(int n_maps, const int *ids, const int64_t *off, const float *pos) {
    std::map<int, std::vector<float>> m;
    for (int k = 0; k < n_maps; k++) m[ids[k]] = std::vector<float>(pos + off[k], pos + off[k + 1]);
    return m;
}
}  // namespace

extern "C" {

float ref_score_f(const float *b, int nbv, const float *x, int nxv, int i, int j, int p, int q,
                  const float *probs, int nprobs) {
    std::vector<float> bv(b, b + nbv), xv(x, x + nxv), pv(probs, probs + nprobs);
    return score_f(bv, xv, i, j, p, q, pv);
}

// maps given in ascending-id order (std::map order).  out gets sum(len-1) floats, same order.
void ref_non_collapse_probs(int n_maps, const int *ids, const int64_t *off, const float *pos, int gen, float *out) {
    auto m = to_map(n_maps, ids, off, pos);
    auto pr = non_collapse_probs(m, gen);
    size_t o = 0;
    for (auto &e : m) {
        auto &v = pr[e.first];
        for (float f : v) out[o++] = f;
    }
}

// returns number of output maps; out_ids/out_off/out_pos sized by caller (2x input).
int ref_make_reverse_cmap(int n_maps, const int *ids, const int64_t *off, const float *pos, int min_map_len,
                          int *out_ids, int64_t *out_off, float *out_pos) {
    auto m = to_map(n_maps, ids, off, pos);
    auto full = make_reverse_cmap(m, min_map_len);
    int k = 0;
    int64_t o = 0;
    out_off[0] = 0;
    for (auto &e : full) {
        out_ids[k] = e.first;
        for (float f : e.second) out_pos[o++] = f;
        out_off[++k] = o;
    }
    return k;
}

// Returns number of maps parsed; fills ids/off/pos up to the given capacities (call with caps=0 to size).
int64_t ref_parse_cmap(const char *fname, int cap_maps, int64_t cap_pos, int *ids, int64_t *off, float *pos,
                       int *n_maps_out) {
    std::map<int, std::vector<float>> m;
    parse_cmap(std::string(fname), m);
    int64_t tot = 0;
    for (auto &e : m) tot += (int64_t)e.second.size();
    *n_maps_out = (int)m.size();
    if ((int)m.size() <= cap_maps && tot <= cap_pos) {
        int k = 0;
        int64_t o = 0;
        off[0] = 0;
        for (auto &e : m) {
            ids[k] = e.first;
            for (float f : e.second) pos[o++] = f;
            off[++k] = o;
        }
    }
    return tot;
}

// init_mode: 0 semiglobal, 1 local, 2 fitting.  b has nb1 = labels+1 entries, x has nx1.
// S_out / prev_out: nb1-1 rows x nx1-1 cols (row nb1-1 of the reference's S is never touched).
// start_mode: 0 sg, 1 local, 2 fixed end.  path_out: up to min(nb,nx) triples (j,q) + scores; returns length.
int ref_dp_align(const float *b, int nb1, const float *x, int nx1, int init_mode, int fitting, int lookback,
                 const int *used, int n_used, const float *xprobs, const float *bprobs, int swap_b_x,
                 int start_mode, float *S_out, int *prev_out, int *start_out, int *path_jq, float *path_s) {
    std::vector<float> bv(b, b + nb1), xv(x, x + nx1);
    std::vector<float> xp(xprobs, xprobs + (nx1 - 1)), bp(bprobs, bprobs + (nb1 - 1));
    std::set<int> usedset(used, used + n_used);
    std::vector<std::vector<float>> S(nb1, std::vector<float>(nx1 - 1));
    std::vector<std::vector<std::array<int, 2>>> previous(nb1, std::vector<std::array<int, 2>>(nx1 - 1));
    if (init_mode == 0) init_semiglobal_aln(S, nx1, nb1, 0);
    else if (init_mode == 1) init_local_aln(S, nx1, nb1, 0);
    else init_fitting_aln(S, nx1, nb1, 0);
    dp_align(S, previous, bv, xv, fitting != 0, lookback, usedset, xp, bp, swap_b_x != 0);
    int nb = nb1 - 1, nx = nx1 - 1;
    if (S_out)
        for (int j = 0; j < nb; j++) memcpy(S_out + (size_t)j * nx, S[j].data(), sizeof(float) * nx);
    if (prev_out)
        for (int j = 0; j < nb; j++)
            for (int q = 0; q < nx; q++) {
                prev_out[2 * ((size_t)j * nx + q)] = previous[j][q][0];
                prev_out[2 * ((size_t)j * nx + q) + 1] = previous[j][q][1];
            }
    std::array<int, 2> st;
    if (start_mode == 0) st = sg_aln_backtrack_start(S, nb1, nx1);
    else if (start_mode == 1) st = local_backtrack_start(S, nb1, nx1);
    else st = {nb - 1, nx - 1};
    start_out[0] = st[0];
    start_out[1] = st[1];
    if (st[0] < 0) return 0;
    auto aln = get_aln_list(S, previous, st);
    for (size_t k = 0; k < aln.size(); k++) {
        path_jq[2 * k] = std::get<0>(aln[k]);
        path_jq[2 * k + 1] = std::get<1>(aln[k]);
        path_s[k] = std::get<2>(aln[k]);
    }
    return (int)aln.size();
}

// Detection-mode scoring of one problem: semiglobal init, dp_align, multi_backtrack_scores(n).
// Returns the number of scores pushed.
int ref_multi_backtrack(const float *b, int nb1, const float *x, int nx1, int lookback, const float *xprobs,
                        const float *bprobs, int swap_b_x, int n, float *scores_out) {
    std::vector<float> bv(b, b + nb1), xv(x, x + nx1);
    std::vector<float> xp(xprobs, xprobs + (nx1 - 1)), bp(bprobs, bprobs + (nb1 - 1));
    std::set<int> usedset;
    std::vector<std::vector<float>> S(nb1, std::vector<float>(nx1 - 1));
    std::vector<std::vector<std::array<int, 2>>> previous(nb1, std::vector<std::array<int, 2>>(nx1 - 1));
    init_semiglobal_aln(S, nx1, nb1, 0);
    dp_align(S, previous, bv, xv, false, lookback, usedset, xp, bp, swap_b_x != 0);
    std::vector<std::tuple<int, int, float>> sv;
    multi_backtrack_scores(S, previous, 0, 0, nb1, nx1, n, sv);
    for (size_t k = 0; k < sv.size(); k++) scores_out[k] = std::get<2>(sv[k]);
    return (int)sv.size();
}

// triples: (seg id, contig id, score) rows; segs/contigs as flattened maps. out: one float per seg (map order).
void ref_compute_score_thresholds(int64_t n_scores, const int *seg_of, const int *contig_of, const float *score,
                                  int n_segs, const int *seg_ids, const int64_t *seg_off, const float *seg_pos,
                                  int n_contigs, const int *contig_ids, const int64_t *contig_off,
                                  const float *contig_pos, double pvalue, int n_detect, float *out) {
    std::vector<std::tuple<int, int, float>> fs;
    for (int64_t k = 0; k < n_scores; k++) fs.emplace_back(seg_of[k], contig_of[k], score[k]);
    auto segs = to_map(n_segs, seg_ids, seg_off, seg_pos);
    auto contigs = to_map(n_contigs, contig_ids, contig_off, contig_pos);
    int saved = n_detect_scores;
    n_detect_scores = n_detect;
    auto thr = compute_score_thresholds(fs, segs, contigs, pvalue);
    n_detect_scores = saved;
    int k = 0;
    for (auto &e : segs) out[k++] = thr[e.first];
}

void ref_partial_threshold(float full_threshold, int first_lab, int last_lab, const float *contig, int nb1,
                           float *out2) {
    std::vector<std::tuple<int, int, float>> aln;
    aln.emplace_back(last_lab, 0, 0.0f);
    aln.emplace_back(first_lab, 0, 0.0f);
    std::vector<float> cv(contig, contig + nb1);
    auto pr = compute_partial_score_threshold(full_threshold, aln, cv);
    out2[0] = pr.first;
    out2[1] = pr.second;
}

// aln given end->start like the reference's aln_list
void ref_aln_stats(int n, const float *scores, float *median, float *minv) {
    std::vector<std::tuple<int, int, float>> aln;
    for (int k = 0; k < n; k++) aln.emplace_back(0, 0, scores[k]);
    *median = compute_aln_median(aln);
    *minv = compute_aln_min(aln);
}

}  // extern "C"
