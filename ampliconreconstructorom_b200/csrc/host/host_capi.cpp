// extern "C" access to the host-side (CPU) preparation helpers of the drop-in executable, so that callers of the
// C ABI (tests, bench.py, other bindings) prepare inputs with the SAME code the executable uses.  These mirror
// SAHelper.cpp / SegAligner.cpp preparation steps; none of them is on the DP hot path.
#include <cstring>
#include <vector>
#include "cmap.h"
#include "segaligner_b200.h"

using namespace sah;

static void to_set(int32_t n_maps, const int32_t *ids, const int64_t *off, const float *pos, MapSet &m) {
    m.ids.assign(n_maps, 0);
    for (int k = 0; k < n_maps; k++) m.ids[(size_t)k] = ids ? ids[k] : k + 1;
    m.off.assign(off, off + n_maps + 1);
    m.pos.assign(pos, pos + off[n_maps]);
}

extern "C" {

int sa_host_non_collapse_probs(int32_t n_maps, const int64_t *off, const float *pos, int32_t inst_gen, float *prob_out) {
    if (n_maps < 0 || !off || !pos || !prob_out) return SA_EINVAL;
    MapSet m;
    to_set(n_maps, nullptr, off, pos, m);
    non_collapse_probs(m, inst_gen);
    if (!m.prob.empty()) memcpy(prob_out, m.prob.data(), sizeof(float) * m.prob.size());
    return SA_OK;
}

int sa_host_make_reverse(int32_t n_maps, const int32_t *ids, const int64_t *off, const float *pos, int32_t min_map_len,
                         int32_t cap_maps, int64_t cap_pos, int32_t *out_ids, int64_t *out_off, float *out_pos,
                         int32_t *n_out, int64_t *n_pos_out) {
    if (n_maps < 0 || !ids || !off || !pos || !n_out || !n_pos_out) return SA_EINVAL;
    MapSet m;
    to_set(n_maps, ids, off, pos, m);
    MapSet r = make_reverse_cmap(m, min_map_len);
    *n_out = r.n_maps();
    *n_pos_out = r.total_size();
    if (!out_ids || !out_off || !out_pos || r.n_maps() > cap_maps || r.total_size() > cap_pos) return SA_OK;  // sizing call
    for (int k = 0; k < r.n_maps(); k++) out_ids[k] = r.ids[(size_t)k];
    memcpy(out_off, r.off.data(), sizeof(int64_t) * r.off.size());
    if (!r.pos.empty()) memcpy(out_pos, r.pos.data(), sizeof(float) * r.pos.size());
    return SA_OK;
}

int sa_host_parse_cmap(const char *path, int32_t cap_maps, int64_t cap_pos, int32_t *out_ids, int64_t *out_off,
                       float *out_pos, int32_t *n_out, int64_t *n_pos_out) {
    if (!path || !n_out || !n_pos_out) return SA_EINVAL;
    MapSet r;
    if (!parse_cmap(path, r)) return SA_EINVAL;
    *n_out = r.n_maps();
    *n_pos_out = r.total_size();
    if (!out_ids || !out_off || !out_pos || r.n_maps() > cap_maps || r.total_size() > cap_pos) return SA_OK;  // sizing call
    for (int k = 0; k < r.n_maps(); k++) out_ids[k] = r.ids[(size_t)k];
    memcpy(out_off, r.off.data(), sizeof(int64_t) * r.off.size());
    if (!r.pos.empty()) memcpy(out_pos, r.pos.data(), sizeof(float) * r.pos.size());
    return SA_OK;
}

int sa_host_score_thresholds(int64_t n_rows, const int32_t *row_seg, const int32_t *row_contig, const float *row_score,
                             int32_t n_segs, const int32_t *seg_ids, const int64_t *seg_off, const float *seg_pos,
                             int32_t n_contigs, const int32_t *contig_ids, const int64_t *contig_off,
                             const float *contig_pos, double pvalue, int32_t n_detect_scores, float *thr_out) {
    if (n_rows < 0 || !thr_out) return SA_EINVAL;
    MapSet segs, contigs;
    to_set(n_segs, seg_ids, seg_off, seg_pos, segs);
    to_set(n_contigs, contig_ids, contig_off, contig_pos, contigs);
    std::vector<ScoreRow> rows((size_t)n_rows);
    for (int64_t k = 0; k < n_rows; k++) rows[(size_t)k] = ScoreRow{row_seg[k], row_contig[k], row_score[k]};
    auto thr = compute_score_thresholds(rows, segs, contigs, pvalue, n_detect_scores);
    for (size_t k = 0; k < thr.size(); k++) thr_out[k] = thr[k].second;
    return SA_OK;
}

int sa_host_format_float(float v, char *buf, int32_t cap) { return format_float(buf, (size_t)cap, v); }

}  // extern "C"
