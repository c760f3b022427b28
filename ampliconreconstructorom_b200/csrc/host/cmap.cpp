// CMAP parsing, map preparation, thresholds and alignment-file output of the drop-in SegAligner executable.
// Behavioural mirror of /root/reference/SegAligner/SAHelper.cpp and of the host-only helpers of
// SegAligner.cpp / main.cpp (file:line cited per function in cmap.h); written around flat arrays.
// fp32 arithmetic here must round exactly like the reference's SSE2 scalar code: this file is compiled with
// -ffp-contract=off and without -march flags.
#include "cmap.h"
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>

namespace sah {

int MapSet::index_of(int id) const {
    auto it = std::lower_bound(ids.begin(), ids.end(), id);
    if (it == ids.end() || *it != id) return -1;
    return (int)(it - ids.begin());
}

static void flatten(const std::map<int, std::vector<float>> &m, MapSet &out) {
    out.ids.clear();
    out.off.assign(1, 0);
    out.pos.clear();
    out.prob.clear();
    size_t tot = 0;
    for (auto &e : m) tot += e.second.size();
    out.pos.reserve(tot);
    for (auto &e : m) {
        out.ids.push_back(e.first);
        out.pos.insert(out.pos.end(), e.second.begin(), e.second.end());
        out.off.push_back((int64_t)out.pos.size());
    }
}

[[noreturn]] static void parse_abort(const char *what, const std::string &tok) {
    // the reference dies with an uncaught std::invalid_argument / std::out_of_range from stoi / stof
    fprintf(stderr, "terminate called after throwing an instance of '%s'\n  what():  %s\n", what, tok.c_str());
    abort();
}

static int ref_stoi(const char *b, const char *e) {
    std::string tok(b, e);
    errno = 0;
    char *endp = nullptr;
    long v = strtol(tok.c_str(), &endp, 10);
    if (endp == tok.c_str()) parse_abort("std::invalid_argument", "stoi");
    if (errno == ERANGE || v < INT32_MIN || v > INT32_MAX) parse_abort("std::out_of_range", "stoi");
    return (int)v;
}

static float ref_stof(const char *b, const char *e) {
    std::string tok(b, e);
    errno = 0;
    char *endp = nullptr;
    float v = strtof(tok.c_str(), &endp);
    if (endp == tok.c_str()) parse_abort("std::invalid_argument", "stof");
    if (errno == ERANGE) parse_abort("std::out_of_range", "stof");
    return v;
}

bool parse_cmap(const std::string &fname, MapSet &out) {
    std::map<int, std::vector<float>> maps;
    FILE *f = fopen(fname.c_str(), "rb");
    if (!f) {
        flatten(maps, out);
        return false;
    }
    std::string data;
    {
        char buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
        fclose(f);
    }
    int curr_id = 0;
    float curr_pos = 0;
    std::vector<float> *cur = nullptr;
    int cur_key = 0;
    bool have_cur = false;
    const char *p = data.data(), *end = p + data.size();
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;  // std::getline: line = [p, le); a final line without '\n' still counts
        if (!(le > p && *p == '#')) {
            int counter = 0;
            const char *s = p;
            for (;;) {
                const char *tab = (const char *)memchr(s, '\t', (size_t)(le - s));
                if (!tab) break;
                if (counter == 0) {
                    const int id = ref_stoi(s, tab);
                    if (curr_id != id) {
                        curr_id = id;
                        maps[curr_id] = std::vector<float>();  // SAHelper.cpp:35: re-created empty
                        have_cur = false;
                    }
                } else if (counter == 5) {
                    curr_pos = ref_stof(s, tab);
                    break;
                }
                s = tab + 1;
                counter++;
            }
            if (!have_cur || cur_key != curr_id) {
                cur = &maps[curr_id];
                cur_key = curr_id;
                have_cur = true;
            }
            cur->push_back(curr_pos);  // SAHelper.cpp:44 (also for short / blank lines: stale position)
        }
        if (!nl) break;
        p = nl + 1;
    }
    flatten(maps, out);
    return true;
}

MapSet make_reverse_cmap(const MapSet &in, int min_map_len) {
    std::map<int, std::vector<float>> full;
    for (int m = 0; m < in.n_maps(); m++) {
        const float *cp = in.map(m);
        const int sz = in.size(m);
        const float map_length = (float)sz;  // SAHelper.cpp:67
        if (map_length <= min_map_len) continue;
        const int id = in.ids[m];
        full[id] = std::vector<float>(cp, cp + sz);
        std::vector<float> rev;
        rev.reserve(sz);
        const float back = cp[sz - 1];
        for (int j = sz - 2; j > -1; --j) rev.push_back(back - cp[j]);
        rev.push_back(back);
        full[-1 * id] = rev;  // for id 0 this overwrites the forward copy, like the reference (:73-78)
    }
    MapSet out;
    flatten(full, out);
    return out;
}

void non_collapse_probs(MapSet &maps, int inst_gen) {
    maps.prob.assign((size_t)std::max<int64_t>(maps.total_size() - maps.n_maps(), 0), 0.0f);
    float non_collapse = 1.0f;  // SegAligner.cpp:43, survives across maps
    for (int m = 0; m < maps.n_maps(); m++) {
        const float *y = maps.map(m);
        const int y_len = maps.labels(m);
        float *out = maps.prob.data() + (maps.off[m] - m);
        float left_prev = 1.0f;
        int o = 0;
        for (int i = 1; i < y_len; ++i) {
            const float dist = y[i] - y[i - 1];
            if (inst_gen == 1) {
                non_collapse = fminf(1.0f, powf(dist, 4) / powf(1950, 4));  // :53
            } else {
                if (dist < 475) non_collapse = 0.0f;
                else non_collapse = (float)fmin(1.0f, (dist - 450) / 1500.0);  // :58 (double)
            }
            out[o++] = left_prev * non_collapse;
            left_prev = non_collapse;
        }
        if (y_len >= 1) out[o++] = left_prev * non_collapse;  // :65
    }
}

std::vector<int> parse_contig_list(const std::string &fname) {
    std::vector<int> v;
    FILE *f = fopen(fname.c_str(), "r");
    if (!f) return v;
    int id;
    while (fscanf(f, "%d", &id) == 1) v.push_back(id);
    fclose(f);
    return v;
}

int format_float(char *buf, size_t cap, float v) { return snprintf(buf, cap, "%g", (double)v); }

bool print_alignment(const std::string &outname, const std::vector<AlnEntry> &aln, int contig_id, int seg_id,
                     int seg_vec_size, float curr_best_score) {
    std::string out;
    out.reserve(256 + aln.size() * 64);
    char num[64];
    const char direction = seg_id < 0 ? '-' : '+';
    const int aseg = seg_id < 0 ? -seg_id : seg_id;
    out += "#seg_seq\ttotal_score\tcircular\n#";
    out += std::to_string(aseg);
    out += direction;
    out += '\t';
    format_float(num, sizeof num, curr_best_score);
    out += num;
    out += "\tFalse\n";
    out += "#contig_id\tseg_id\tcontig_label\tseg_label\tcontig_dir\tseg_dir\tseg_aln_number\tscore\tscore_delta\n";
    float prev_score = 0.0f;
    for (int i = (int)aln.size() - 1; i > -1; --i) {
        const float curr_score = aln[i].s;
        const float score_delta = curr_score - prev_score;
        int curr_lab = aln[i].q;
        if (seg_id < 0) curr_lab = seg_vec_size - curr_lab - 1;  // SAHelper.cpp:119-123
        else curr_lab += 1;
        out += std::to_string(contig_id);
        out += '\t';
        out += std::to_string(aseg);
        out += '\t';
        out += std::to_string(aln[i].j + 1);
        out += '\t';
        out += std::to_string(curr_lab);
        out += "\t+\t";
        out += direction;
        out += "\t0\t";
        format_float(num, sizeof num, curr_score);
        out += num;
        out += '\t';
        format_float(num, sizeof num, score_delta);
        out += num;
        out += '\n';
        prev_score = curr_score;
    }
    FILE *f = fopen(outname.c_str(), "wb");
    if (!f) return false;
    fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    return true;
}

// number of host threads for the embarrassingly parallel text / sort phases (not the reference's -nthreads, which
// only determines the row order of *_all_scores.txt)
static int host_threads(size_t work_items) {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    const size_t by_work = work_items / 65536 + 1;  // below ~64k items a thread is not worth starting
    return (int)std::max<size_t>(1, std::min<size_t>(std::min<unsigned>(hw, 32u), by_work));
}

bool write_all_scores(const std::vector<ScoreRow> &rows, const std::string &prefix) {
    FILE *f = fopen((prefix + "_all_scores.txt").c_str(), "wb");
    if (!f) return false;
    fputs("#seg\tcontig\tscore\n", f);
    // format blocks of rows in parallel, write them in order
    const size_t block = 1u << 18;
    const int nt = host_threads(rows.size());
    for (size_t base = 0; base < rows.size(); base += block * (size_t)nt) {
        std::vector<std::string> out((size_t)nt);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) {
            const size_t lo = std::min(rows.size(), base + block * (size_t)t), hi = std::min(rows.size(), lo + block);
            if (lo >= hi) break;
            auto work = [&rows, &out, t, lo, hi]() {
                std::string &o = out[(size_t)t];
                o.reserve((hi - lo) * 24);
                char num[64];
                for (size_t k = lo; k < hi; k++) {
                    const ScoreRow &r = rows[k];
                    int n = snprintf(num, sizeof num, "%d\t%d\t", r.seg, r.contig);
                    o.append(num, (size_t)n);
                    n = format_float(num, sizeof num, r.score);
                    o.append(num, (size_t)n);
                    o += '\n';
                }
            };
            if (nt == 1) work();
            else th.emplace_back(work);
        }
        for (auto &t : th) t.join();
        for (auto &o : out) fwrite(o.data(), 1, o.size(), f);
    }
    fclose(f);
    return true;
}

bool write_score_thresholds(const std::vector<std::pair<int, float>> &thr, const std::string &full_prefix) {
    FILE *f = fopen((full_prefix + "_score_thresholds.txt").c_str(), "wb");
    if (!f) return false;
    fputs("#seg\tscore\n", f);
    char num[64];
    for (auto &e : thr) {
        format_float(num, sizeof num, e.second);
        fprintf(f, "%d\t%s\n", e.first, num);
    }
    fclose(f);
    return true;
}

std::vector<std::pair<int, float>> compute_score_thresholds(const std::vector<ScoreRow> &rows, const MapSet &segs,
                                                            const MapSet &contigs, double pvalue, int n_detect_scores) {
    std::vector<std::vector<std::pair<int, float>>> out;
    compute_score_thresholds_multi(rows, segs, contigs, std::vector<double>{pvalue}, n_detect_scores, out);
    return out[0];
}

void compute_score_thresholds_multi(const std::vector<ScoreRow> &rows, const MapSet &segs, const MapSet &contigs,
                                    const std::vector<double> &pvalues, int n_detect_scores,
                                    std::vector<std::vector<std::pair<int, float>>> &out) {
    const int ns = segs.n_maps();
    std::vector<std::vector<float>> seg_scores((size_t)ns);
    {
        // rows arrive in long runs of the same segment id: remember the last lookup
        std::vector<int32_t> seg_of(rows.size());
        std::vector<size_t> cnt((size_t)ns, 0);
        int last_id = 0, last_k = -2;
        for (size_t r = 0; r < rows.size(); r++) {
            if (last_k == -2 || rows[r].seg != last_id) {
                last_id = rows[r].seg;
                last_k = segs.index_of(last_id);
            }
            seg_of[r] = last_k;
            if (last_k >= 0) cnt[(size_t)last_k]++;
        }
        for (int k = 0; k < ns; k++) seg_scores[(size_t)k].reserve(cnt[(size_t)k]);
        for (size_t r = 0; r < rows.size(); r++)
            if (seg_of[r] >= 0) seg_scores[(size_t)seg_of[r]].push_back(rows[r].score);
    }
    const int right_cutoff = 25;
    const int left_cut = (int)round(0.15 * fmin((double)contigs.n_maps(), (double)n_detect_scores));  // main.cpp:66
    int m = 0;
    for (int c = 0; c < contigs.n_maps(); c++) m += contigs.labels(c);  // :69-72
    out.assign(pvalues.size(), std::vector<std::pair<int, float>>((size_t)ns));
    auto one_segment = [&](int k) {
        std::vector<float> &sc = seg_scores[(size_t)k];
        const int n = segs.labels(k);  // :76
        const int64_t sz = (int64_t)sc.size();
        const int64_t first = sz - (right_cutoff + left_cut);
        int64_t last = sz - 1 - right_cutoff;
        if (n < 12) last -= 25;  // :80-82
        const bool bad = first < 0 || last > sz || last < first;
        if (!bad) {
            // the reference sorts the whole vector (:59) but only reads [first, last): order that tail only
            std::nth_element(sc.begin(), sc.begin() + first, sc.end());
            std::sort(sc.begin() + first, sc.end());
        }
        for (size_t pv = 0; pv < pvalues.size(); pv++) {
            const double E_cutoff = -log(1 - pvalues[pv]);
            float thr;
            if (bad) {
                thr = -NAN;  // reference: iterators outside the vector (UB); observed output "-nan"
            } else {
                const int64_t cnt = last - first;
                const double nn = (double)cnt;
                double sx = 0.0, sy = 0.0;
                for (int64_t t = 0; t < cnt; t++) sx += sc[(size_t)(first + t)];
                for (int64_t t = 0; t < cnt; t++) sy += (int)log((double)(cnt - t));  // int-truncated ln(rank), :85-90
                const double xbar = sx / nn, ybar = sy / nn;  // linreg, SegAligner.cpp:304-319
                double cov = 0.0, var = 0.0;
                for (int64_t t = 0; t < cnt; t++) {
                    const double xv = sc[(size_t)(first + t)];
                    const double yv = (int)log((double)(cnt - t));
                    cov += (xv - xbar) * (yv - ybar);
                    var += (xv - xbar) * (xv - xbar);
                }
                const double slope = cov / var;
                const double intercept = ybar - slope * xbar;
                const int mn = (int)((unsigned)m * (unsigned)n);  // int product of the reference (:97), wraps
                const double K = exp(intercept - log((double)mn));
                const double S_cutoff = -log(E_cutoff / (K * m * n)) / (-slope);  // :100
                thr = (float)S_cutoff;
            }
            out[pv][(size_t)k] = {segs.ids[(size_t)k], thr};
        }
    };
    const int nt = host_threads(rows.size());
    if (nt == 1) {
        for (int k = 0; k < ns; k++) one_segment(k);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&, t]() {
                for (int k = t; k < ns; k += nt) one_segment(k);
            });
        for (auto &t : th) t.join();
    }
}

void partial_score_threshold(float full_threshold, const std::vector<AlnEntry> &aln, const float *contig, int contig_size,
                             float &by_label, float &by_length) {
    const int first_lab = aln.back().j;
    const int last_lab = aln[0].j;
    const int aln_span_labs = last_lab - first_lab + 1;
    const float tot_labs = (float)(contig_size - 1);
    by_label = full_threshold * (aln_span_labs / tot_labs);
    const float aln_span_bases = contig[last_lab] - contig[first_lab];
    by_length = full_threshold * (aln_span_bases / contig[contig_size - 1]);
}

static void score_deltas(const std::vector<AlnEntry> &aln, std::vector<float> &d) {
    d.clear();
    float prev = 0.0f;
    for (int i = (int)aln.size() - 2; i > -1; i--) {
        d.push_back(aln[i].s - prev);
        prev = aln[i].s;
    }
}

float aln_median(const std::vector<AlnEntry> &aln) {
    std::vector<float> d;
    score_deltas(aln, d);
    std::sort(d.begin(), d.end());
    if (d.empty()) return NAN;  // reference: reads d[0] and d[-1] of an empty vector (UB); unreachable via main
    const int mid = (int)d.size() / 2;
    if (d.size() % 2) return d[(size_t)mid];
    return (d[(size_t)mid] + d[(size_t)mid - 1]) / 2;
}

float aln_min(const std::vector<AlnEntry> &aln) {
    std::vector<float> d;
    score_deltas(aln, d);
    if (d.empty()) return NAN;
    return *std::min_element(d.begin(), d.end());
}

}  // namespace sah
