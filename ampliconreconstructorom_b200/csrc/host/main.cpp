// SegAligner -- B200-native drop-in for /root/reference/SegAligner (main.cpp:453-656).
//
// Same positional arguments, same flags (parse_args, main.cpp:363-435), same output files, same stdout
// progress lines; the three phase drivers run_SA_score / run_SA_aln / run_SA_fitting (main.cpp:112-358) are
// re-designed around batched GPU calls through the C ABI of include/segaligner_b200.h:
//   phase 1  every (segment+-, contig) problem of the run in ONE sa_score_batch / sa_detect_batch per GPU;
//   phase 3/4 the reference's per-pair `while` loop (main.cpp:268) becomes rounds: round r aligns every pair
//            that is still live in one sa_align_batch, then the host applies thresholds and filters.
// -nthreads is accepted; it only fixes the ROW ORDER of *_all_scores.txt (the reference concatenates its
// per-thread result vectors, main.cpp:547-551), which we reproduce.  Extra flag: -ngpus=N (default: all
// visible GPUs for big runs, 1 otherwise; unknown flags are ignored by the reference's parser, so callers
// written for the reference are unaffected).  There is no CPU fallback.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <unistd.h>
#include "cmap.h"
#include "segaligner_b200.h"

using namespace sah;

namespace {

// Leaving a job early: the stand-alone process turns this into exit(code); the resident server replies with the
// code and keeps running (unless fatal: a CUDA failure leaves the context unusable).
struct JobExit {
    int code;
    bool fatal;
};

// ---- globals of the reference (main.cpp:19-37) -------------------------------------------------------------
struct Options {
    bool local_aln = false, detection = false, fitting_aln = false, swap_b_x = false;
    int n_detect_scores = 500, n_threads = 1, min_map_len = 10, lookback = 6, inst_gen = 2;
    int max_seg_contig_alns = 6;
    float p_val = 0.0001f, p_val_tip = 0.000001f, p_val_RG = 0.000000001f;
    std::string sample_prefix = "SA_output", detection_label, contig_list_file;
    bool limit_lookback = true, do_tip = true;
    int n_gpus = 0;  // 0 = auto
};

bool starts_with(const std::string &s, const char *p) { return s.rfind(p, 0) == 0; }
std::string after_eq(const std::string &s) { return s.substr(s.find('=') + 1); }

int ref_stoi_arg(const std::string &v) {
    char *e = nullptr;
    long r = strtol(v.c_str(), &e, 10);
    if (e == v.c_str()) {
        fprintf(stderr, "terminate called after throwing an instance of 'std::invalid_argument'\n  what():  stoi\n");
        abort();
    }
    return (int)r;
}

void parse_args(int argc, char **argv, Options &o) {
    for (int i = 3; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-nl") o.limit_lookback = false;
        else if (a == "-no_tip_aln") o.do_tip = false;
        else if (a == "-detection") {
            o.local_aln = o.detection = o.swap_b_x = true;
            o.do_tip = false;
            o.detection_label = "_detection";
            o.p_val = o.p_val_RG;
            printf("detection mode ON, local ON, swap_b_x ON\n");
            fflush(stdout);
        } else if (starts_with(a, "-prefix=")) o.sample_prefix = after_eq(a);
        else if (starts_with(a, "-contig_list=")) o.contig_list_file = after_eq(a);
        else if (a == "-local") {
            o.local_aln = true;
            printf("Local alignment mode\n");
            fflush(stdout);
        } else if (a == "-fitting") {
            o.fitting_aln = true;
            printf("Fitting alignment mode. Multithreading will be off.\n\n");
            fflush(stdout);
        } else if (starts_with(a, "-nthreads=")) o.n_threads = ref_stoi_arg(after_eq(a));
        else if (starts_with(a, "-min_labs=")) o.min_map_len = ref_stoi_arg(after_eq(a));
        else if (starts_with(a, "-n_detect_scores=")) o.n_detect_scores = ref_stoi_arg(after_eq(a));
        else if (starts_with(a, "-gen=")) {
            o.inst_gen = ref_stoi_arg(after_eq(a));
            if (o.inst_gen != 1 && o.inst_gen != 2) {
                printf("1 or 2 are the only valid arguments for -gen");
                throw JobExit{1, false};
            }
        } else if (starts_with(a, "-ngpus=")) o.n_gpus = ref_stoi_arg(after_eq(a));
    }
}

// ---- multi-GPU engine: one context + one host thread per device, problems dealt largest-first ----------------
struct Engine {
    std::vector<sa_ctx *> ctx;
    const MapSet *contigs = nullptr, *segs = nullptr;

    void die(sa_ctx *c, const char *what, int rc) {
        fprintf(stderr, "SegAligner(b200): %s failed (%d): %s\n", what, rc, c ? sa_last_error(c) : "");
        throw JobExit{2, true};
    }
    // SA_GPU_OVERSUBSCRIBE=1 (testing aid): -ngpus=N may exceed the visible devices; context g then lives on device
    // g % visible, so the sharded multi-GPU path (shard / score / align below) can be exercised on a one-GPU box.
    static bool oversubscribe() {
        const char *e = getenv("SA_GPU_OVERSUBSCRIBE");
        return e && atoi(e) != 0;
    }
    void open_range(int have, int n, int avail) {
        ctx.resize((size_t)n, nullptr);
        std::vector<std::thread> th;
        std::vector<int> rcs((size_t)n, 0);
        for (int g = have; g < n; g++) th.emplace_back([&, g]() { rcs[(size_t)g] = sa_ctx_create(g % avail, &ctx[(size_t)g]); });
        for (auto &t : th) t.join();
        for (int g = have; g < n; g++)
            if (rcs[(size_t)g]) die(nullptr, "sa_ctx_create (need an sm_100 GPU)", rcs[(size_t)g]);
    }
    void open(int n) {
        if (!ctx.empty()) return;  // resident server: contexts stay open between jobs
        const int avail = sa_device_count();
        if (avail <= 0) {
            fprintf(stderr, "SegAligner(b200): no CUDA device visible; this build has no CPU fallback\n");
            throw JobExit{2, true};
        }
        n = std::max(1, oversubscribe() ? std::min(n, 64) : std::min(n, avail));
        open_range(0, n, avail);
    }
    void open_more(int n) {  // device 0 is already open; add devices 1..n-1
        const int avail = sa_device_count();
        n = std::max(1, oversubscribe() ? std::min(n, 64) : std::min(n, avail));
        const int have = (int)ctx.size();
        if (n <= have) return;
        open_range(have, n, avail);
    }
    void close() {
        for (auto c : ctx) sa_ctx_destroy(c);
        ctx.clear();
    }
    void set_maps(const MapSet &c, const MapSet &s) {
        contigs = &c;
        segs = &s;
        parallel([&](int g) {
            int rc = sa_set_maps(ctx[(size_t)g], SA_SIDE_CONTIG, c.n_maps(), c.off.data(), c.pos.data(), c.prob.data());
            if (rc) die(ctx[(size_t)g], "sa_set_maps(contigs)", rc);
            rc = sa_set_maps(ctx[(size_t)g], SA_SIDE_SEG, s.n_maps(), s.off.data(), s.pos.data(), s.prob.data());
            if (rc) die(ctx[(size_t)g], "sa_set_maps(segments)", rc);
        });
    }
    // f(g) on one host thread per GPU.  die() throws JobExit; an exception must not escape a std::thread (that would be
    // std::terminate instead of the documented exit code 2), so each worker parks its JobExit and the first one is
    // rethrown on the calling thread after the join.
    template <class F>
    void parallel(F f) {
        if (ctx.size() == 1) {
            f(0);
            return;
        }
        std::vector<std::thread> th;
        std::vector<JobExit> errs(ctx.size(), JobExit{0, false});
        for (int g = 0; g < (int)ctx.size(); g++)
            th.emplace_back([&, g]() {
                try {
                    f(g);
                } catch (const JobExit &e) {
                    errs[(size_t)g] = e.code ? e : JobExit{2, true};
                }
            });
        for (auto &t : th) t.join();
        for (const JobExit &e : errs)
            if (e.code) throw e;
    }
    // shard[g] = indices of the problems GPU g handles.  Problems are dealt round-robin WITHIN size classes
    // (1/4-octave of cells), largest classes first, so every GPU gets the same mix of sizes in O(n) (the library
    // orders each shard largest-first itself); a full sort of 10^7 problems would cost more than their scoring.
    std::vector<std::vector<int64_t>> shard(const std::vector<sa_problem> &pr) const {
        const int G = (int)ctx.size();
        std::vector<std::vector<int64_t>> sh((size_t)G);
        if (G == 1) {
            sh[0].resize(pr.size());
            for (size_t k = 0; k < pr.size(); k++) sh[0][k] = (int64_t)k;
            return sh;
        }
        const int NB = 4 * 64;
        std::vector<uint8_t> cls(pr.size());
        std::vector<int64_t> cnt((size_t)NB + 1, 0);
        for (size_t k = 0; k < pr.size(); k++) {
            const double cells = (double)contigs->labels(pr[k].contig) * (double)segs->labels(pr[k].seg);
            int b = cells > 1.0 ? (int)(std::log2(cells) * 4.0) : 0;
            b = NB - 1 - std::min(b, NB - 1);  // descending size
            cls[k] = (uint8_t)b;
            cnt[(size_t)b + 1]++;
        }
        for (int b = 0; b < NB; b++) cnt[(size_t)b + 1] += cnt[(size_t)b];
        std::vector<int64_t> rank(pr.size());  // position in the size-class order
        for (size_t k = 0; k < pr.size(); k++) rank[k] = cnt[cls[k]]++;
        for (auto &v : sh) v.reserve(pr.size() / (size_t)G + 1);
        std::vector<int64_t> by_rank(pr.size());
        for (size_t k = 0; k < pr.size(); k++) by_rank[(size_t)rank[k]] = (int64_t)k;
        for (size_t r = 0; r < by_rank.size(); r++) sh[r % (size_t)G].push_back(by_rank[r]);
        return sh;
    }

    void score(const sa_mode &mode, const std::vector<sa_problem> &pr, std::vector<sa_result> &res) {
        res.assign(pr.size(), sa_result{0, -1, -1, 0});
        if (pr.empty()) return;
        auto sh = shard(pr);
        parallel([&](int g) {
            const auto &idx = sh[(size_t)g];
            if (idx.empty()) return;
            std::vector<sa_problem> p(idx.size());
            std::vector<sa_result> r(idx.size());
            for (size_t k = 0; k < idx.size(); k++) p[k] = pr[(size_t)idx[k]];
            int rc = sa_score_batch(ctx[(size_t)g], &mode, p.data(), (int64_t)p.size(), r.data());
            if (rc) die(ctx[(size_t)g], "sa_score_batch", rc);
            for (size_t k = 0; k < idx.size(); k++) res[(size_t)idx[k]] = r[k];
        });
    }

    // detection scoring: out[k] = the scores multi_backtrack_scores harvests for problem k, in order
    void detect(const sa_mode &mode, const std::vector<sa_problem> &pr, const std::vector<int32_t> &n_want,
                std::vector<std::vector<float>> &out) {
        out.assign(pr.size(), {});
        if (pr.empty()) return;
        auto sh = shard(pr);
        parallel([&](int g) {
            const auto &idx = sh[(size_t)g];
            if (idx.empty()) return;
            const size_t m = idx.size();
            std::vector<sa_problem> p(m);
            std::vector<int32_t> want(m), got(m);
            std::vector<int64_t> off(m + 1, 0);
            for (size_t k = 0; k < m; k++) {
                p[k] = pr[(size_t)idx[k]];
                want[k] = n_want[(size_t)idx[k]];
                off[k + 1] = off[k] + want[k];
            }
            std::vector<float> sc((size_t)std::max<int64_t>(off[m], 1));
            int rc = sa_detect_batch(ctx[(size_t)g], &mode, p.data(), (int64_t)m, want.data(), off.data(), sc.data(), got.data());
            if (rc) die(ctx[(size_t)g], "sa_detect_batch", rc);
            for (size_t k = 0; k < m; k++) out[(size_t)idx[k]].assign(sc.begin() + off[k], sc.begin() + off[k] + got[k]);
        });
    }

    // alignment: out.get(k) = result and traceback (end->start) of problem k, read in place from the pinned buffers
    // of the context that computed it (valid until the next align call).
    struct AlignOut {
        struct View {
            const sa_result *res;
            const int32_t *pj, *pq;
            const float *ps;
        };
        std::vector<sa_packed_paths> pk;        // per GPU
        std::vector<std::vector<int64_t>> idx;  // per GPU: the problems it handled (empty for a single GPU)
        std::vector<int32_t> gpu_of;            // per problem (multi-GPU only)
        std::vector<int64_t> local_of;
        View get(size_t k) const {
            const size_t g = gpu_of.empty() ? 0 : (size_t)gpu_of[k];
            const size_t l = gpu_of.empty() ? k : (size_t)local_of[k];
            const sa_packed_paths &p = pk[g];
            return View{p.results + l, p.path_j + p.start[l], p.path_q + p.start[l], p.path_s + p.start[l]};
        }
    };
    // pr[k].used_slot = WORD offset of the problem's skip bitmap inside `arena` (or -1): the C ABI is used with one
    // word per slot, so bitmaps of different lengths sit packed and the arena is uploaded as it is.
    void align(const sa_mode &mode, std::vector<sa_problem> &pr, const std::vector<uint32_t> &arena, AlignOut &out) {
        const size_t n = pr.size();
        const size_t G = ctx.size();
        out.pk.assign(G, sa_packed_paths{});
        out.idx.clear();
        out.gpu_of.clear();
        out.local_of.clear();
        if (n == 0) return;
        const uint32_t *bits = arena.empty() ? nullptr : arena.data();
        const int64_t n_words = (int64_t)arena.size();
        if (G == 1) {
            int rc = sa_align_batch_packed(ctx[0], &mode, pr.data(), (int64_t)n, bits, 1, n_words, &out.pk[0]);
            if (rc) die(ctx[0], "sa_align_batch_packed", rc);
            return;
        }
        out.idx = shard(pr);
        out.gpu_of.resize(n);
        out.local_of.resize(n);
        parallel([&](int g) {
            const auto &idx = out.idx[(size_t)g];
            if (idx.empty()) return;
            std::vector<sa_problem> p(idx.size());
            for (size_t k = 0; k < idx.size(); k++) {
                p[k] = pr[(size_t)idx[k]];
                out.gpu_of[(size_t)idx[k]] = g;
                out.local_of[(size_t)idx[k]] = (int64_t)k;
            }
            int rc = sa_align_batch_packed(ctx[(size_t)g], &mode, p.data(), (int64_t)p.size(), bits, 1, n_words, &out.pk[(size_t)g]);
            if (rc) die(ctx[(size_t)g], "sa_align_batch_packed", rc);
        });
    }

    // Device-resident alignment loop (sa_align_rounds): every round of every pair in one call per GPU.  out.nr(k) = number
    // of rounds of pair k (-1: the pair did not fit the device scratch, the caller runs it through align()).
    struct RoundsOut {
        std::vector<sa_rounds_out> ro;  // per GPU
        std::vector<int32_t> gpu_of;    // per pair (multi-GPU only)
        std::vector<int64_t> local_of;
        int max_alns = 0;
        int nr(size_t k) const {
            const size_t g = gpu_of.empty() ? 0 : (size_t)gpu_of[k];
            const size_t l = gpu_of.empty() ? k : (size_t)local_of[k];
            return ro[g].n_rounds[l];
        }
        struct View {
            const sa_round *rec;
            const int32_t *pj, *pq;
            const float *ps;
        };
        View get(size_t k, int r) const {
            const size_t g = gpu_of.empty() ? 0 : (size_t)gpu_of[k];
            const size_t l = gpu_of.empty() ? k : (size_t)local_of[k];
            const sa_rounds_out &o = ro[g];
            const sa_round *rec = o.rounds + l * (size_t)max_alns + (size_t)r;
            return View{rec, o.path_j + rec->path_off, o.path_q + rec->path_off, o.path_s + rec->path_off};
        }
    };
    void align_rounds(const sa_mode &mode, const std::vector<sa_problem> &pr, const std::vector<float> &thr, int max_alns,
                      const std::vector<uint32_t> &arena, RoundsOut &out) {
        const size_t n = pr.size(), G = ctx.size();
        out.ro.assign(G, sa_rounds_out{});
        out.gpu_of.clear();
        out.local_of.clear();
        out.max_alns = max_alns;
        if (n == 0) return;
        const uint32_t *bits = arena.empty() ? nullptr : arena.data();
        const int64_t n_words = (int64_t)arena.size();
        if (G == 1) {
            int rc = sa_align_rounds(ctx[0], &mode, pr.data(), (int64_t)n, thr.data(), max_alns, bits, 1, n_words, &out.ro[0]);
            if (rc) die(ctx[0], "sa_align_rounds", rc);
            return;
        }
        auto sh = shard(pr);
        out.gpu_of.resize(n);
        out.local_of.resize(n);
        parallel([&](int g) {
            const auto &idx = sh[(size_t)g];
            if (idx.empty()) return;
            std::vector<sa_problem> p(idx.size());
            std::vector<float> t(idx.size());
            for (size_t k = 0; k < idx.size(); k++) {
                p[k] = pr[(size_t)idx[k]];
                t[k] = thr[(size_t)idx[k]];
                out.gpu_of[(size_t)idx[k]] = g;
                out.local_of[(size_t)idx[k]] = (int64_t)k;
            }
            int rc = sa_align_rounds(ctx[(size_t)g], &mode, p.data(), (int64_t)p.size(), t.data(), max_alns, bits, 1, n_words, &out.ro[(size_t)g]);
            if (rc) die(ctx[(size_t)g], "sa_align_rounds", rc);
        });
    }
};

using Bitmap = std::vector<uint32_t>;
inline size_t bm_words(int labels) { return (size_t)std::max((labels + 31) / 32, 1); }
inline void bm_set(Bitmap &b, int j) { b[(size_t)j >> 5] |= 1u << (j & 31); }
inline bool bm_any(const Bitmap &b) {
    for (uint32_t w : b)
        if (w) return true;
    return false;
}

// One pass of run_SA_aln (main.cpp:213-358) over a set of (segment, contig) pairs, as batched rounds.
// used_seed: per contig the labels already taken (tip phase), else nullptr.  Returns per contig the labels
// of the alignments that were WRITTEN (discovered_contig_used_labels).
double wall_now();
std::vector<Bitmap> run_alignment_pass(Engine &eng, const Options &o, const MapSet &segs, const MapSet &contigs,
                                       const std::vector<std::pair<int, int>> &pairs /* (seg idx, contig idx) */,
                                       const std::vector<std::pair<int, float>> &thresholds /* by seg idx */,
                                       const std::vector<Bitmap> *used_seed, bool tip_aln) {
    const int min_aln_labels = 6;
    const int min_tip_aln_labels = (o.inst_gen == 1) ? 3 : 5;
    std::vector<Bitmap> discovered((size_t)contigs.n_maps());
    for (int c = 0; c < contigs.n_maps(); c++) discovered[(size_t)c].assign(bm_words(contigs.labels(c)), 0u);
    const float med_thresh = tip_aln ? 8750.0f : 8000.0f;
    const float mean_thresh = tip_aln ? 8750.0f : 7000.0f;
    int max_alns = o.max_seg_contig_alns;
    if (tip_aln && o.inst_gen == 2) max_alns = 6;

    // Skip bitmaps live packed in one arena of 32-bit words that is uploaded as it is: first the seed bitmap of every
    // contig that has one (tip phase; shared by all pairs of the contig until they accept an alignment), then one
    // private copy per pair, appended when the pair accepts its first alignment.  10^7 pairs stay cheap.
    std::vector<uint32_t> arena;
    std::vector<int64_t> seed_off((size_t)contigs.n_maps(), -1);
    if (used_seed)
        for (int c = 0; c < contigs.n_maps(); c++)
            if (bm_any((*used_seed)[(size_t)c])) {
                seed_off[(size_t)c] = (int64_t)arena.size();
                arena.insert(arena.end(), (*used_seed)[(size_t)c].begin(), (*used_seed)[(size_t)c].end());
            }
    struct PairState {
        int s, c;
        float thr, curr_best, exp_thresh;
        int count;
        int64_t used_off;  // word offset of the skip bitmap in the arena, -1 = none
        bool own, live;
    };
    std::vector<PairState> st(pairs.size());
    for (size_t k = 0; k < pairs.size(); k++) {
        PairState &p = st[k];
        p.s = pairs[k].first;
        p.c = pairs[k].second;
        p.thr = thresholds[(size_t)p.s].second;
        p.curr_best = p.thr;
        p.exp_thresh = p.thr;
        p.count = 0;
        p.used_off = seed_off[(size_t)p.c];
        p.own = false;
        p.live = true;
    }
    sa_mode mode;
    mode.init_mode = o.local_aln ? SA_INIT_LOCAL : SA_INIT_SEMIGLOBAL;
    mode.start_mode = o.local_aln ? SA_START_LOCAL : SA_START_SEMIGLOBAL;
    mode.fitting = 0;
    mode.lookback = o.lookback;
    mode.swap_b_x = o.swap_b_x ? 1 : 0;

    // compute_partial_score_threshold (SegAligner.cpp:23-36) on a flat path (end -> start): max of the two, main.cpp:290
    auto partial_thresh = [&](const PairState &p, int len, const int32_t *pj) -> float {
        const float *contig = contigs.map(p.c);
        const int contig_size = contigs.size(p.c);
        const int first_lab = pj[len - 1], last_lab = pj[0];
        const int aln_span_labs = last_lab - first_lab + 1;
        const float tot_labs = (float)(contig_size - 1);
        const float by_lab = p.thr * (aln_span_labs / tot_labs);
        const float aln_span_bases = contig[last_lab] - contig[first_lab];
        const float by_len = p.thr * (aln_span_bases / contig[contig_size - 1]);
        return fmaxf(by_lab, by_len);
    };
    // An ACCEPTED alignment (p.count already incremented): the tip / length / score filters of main.cpp:303-325, the file,
    // and the labels of the written alignment (SAHelper.cpp:115).
    auto emit = [&](const PairState &p, int len, const int32_t *pj, const int32_t *pq, const float *ps, std::vector<AlnEntry> &aln,
                    std::vector<std::pair<int, int>> &disc) {
        if (tip_aln) {  // must touch a contig end (:303-308)
            const size_t last_lab_gap = (size_t)(contigs.size(p.c) - 2) - (size_t)pj[0];
            if (last_lab_gap > 3 && pj[len - 1] > 2) return;
        }
        if (len < min_tip_aln_labels || (!tip_aln && len < min_aln_labels)) return;
        aln.resize((size_t)len);
        for (int t = 0; t < len; t++) aln[(size_t)t] = AlnEntry{pj[t], pq[t], ps[t]};
        const float mean = aln[0].s / (float)(aln.size() - 1);
        const float median = aln_median(aln);
        if (aln.size() < 5) {
            const float min_score = aln_min(aln);
            if (mean < 9400 || median < 9400 || min_score < 9200) return;
        } else if (mean < mean_thresh || median < med_thresh) {
            return;
        }
        const int x = segs.ids[(size_t)p.s];
        const int contig_id = contigs.ids[(size_t)p.c];
        std::string seg_id = std::to_string(x < 0 ? -x : x);
        if (x < 0) seg_id += "_r";
        const std::string outname = o.sample_prefix + "_" + std::to_string(contig_id) + "_" + seg_id + "_" + std::to_string(p.count) +
                                    (tip_aln ? "_tip" : "") + "_aln.txt";
        print_alignment(outname, aln, contig_id, x, segs.size(p.s), p.curr_best);
        for (auto &e : aln) disc.emplace_back(p.c, e.j);
    };

    // ---- device-resident loop: every round of every pair in ONE call per GPU (sa_align_rounds); the host replays the
    // bookkeeping on the returned rounds to apply the filters and write the files.  Semi-global, banded, unswapped runs only
    // (-local / -detection / -nl keep the batched rounds below); pairs too large for the device scratch come back as -1.
    bool use_rounds = !o.local_aln && !o.swap_b_x && o.lookback == 6 && !st.empty();
    if (const char *e = getenv("SA_NO_ROUNDS")) use_rounds = use_rounds && atoi(e) == 0;
    for (int sidx = 0; use_rounds && sidx < segs.n_maps(); sidx++) use_rounds = segs.labels(sidx) >= 2;
    if (use_rounds) {
        std::vector<sa_problem> rp(st.size());
        std::vector<float> rthr(st.size());
        for (size_t k = 0; k < st.size(); k++) {
            rp[k] = sa_problem{st[k].c, st[k].s, (int32_t)st[k].used_off, 0};
            rthr[k] = st[k].thr;
        }
        const double t_r0 = wall_now();
        Engine::RoundsOut ro;
        eng.align_rounds(mode, rp, rthr, max_alns, arena, ro);
        const double t_r1 = wall_now();
        const size_t n_pairs = st.size();
        const int n_workers = (int)std::max<size_t>(1, std::min<size_t>(std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u),
                                                                        n_pairs / 50000 + 1));
        std::vector<std::vector<std::pair<int, int>>> disc((size_t)n_workers);
        std::vector<int> bad((size_t)n_workers, 0);
        size_t n_fallback = 0;
        auto work = [&](int wi) {
            std::vector<AlnEntry> aln;
            const size_t lo = n_pairs * (size_t)wi / (size_t)n_workers, hi = n_pairs * ((size_t)wi + 1) / (size_t)n_workers;
            for (size_t k = lo; k < hi; k++) {
                PairState &p = st[k];
                const int nr = ro.nr(k);
                if (nr < 0) continue;  // left to the batched rounds below
                p.live = false;
                bool more = p.curr_best >= p.exp_thresh && p.count < max_alns;  // main.cpp:268 (false for a NaN threshold)
                int r = 0;
                for (; more && r < nr; r++) {
                    const Engine::RoundsOut::View v = ro.get(k, r);
                    const int len = v.rec->path_len;
                    if (len <= 0) {  // no backtrack start (cannot happen with >= 1 label each): the device stopped here too
                        more = false;
                        r++;
                        break;
                    }
                    p.exp_thresh = partial_thresh(p, len, v.pj);
                    p.curr_best = v.rec->score;
                    if (!(p.curr_best > p.exp_thresh)) {  // not accepted: the loop ends (equality would spin in the reference)
                        more = false;
                        r++;
                        break;
                    }
                    p.count += 1;
                    emit(p, len, v.pj, v.pq, v.ps, aln, disc[(size_t)wi]);
                    more = p.count < max_alns;
                }
                if (more || r != nr) bad[(size_t)wi]++;  // the device loop and this replay must agree on the number of rounds
            }
        };
        if (n_workers == 1) {
            work(0);
        } else {
            std::vector<std::thread> th;
            for (int wi = 0; wi < n_workers; wi++) th.emplace_back(work, wi);
            for (auto &t : th) t.join();
        }
        int n_bad = 0;
        for (int wi = 0; wi < n_workers; wi++) {
            n_bad += bad[(size_t)wi];
            for (auto &e : disc[(size_t)wi]) bm_set(discovered[(size_t)e.first], e.second);
        }
        for (size_t k = 0; k < n_pairs; k++) n_fallback += st[k].live ? 1 : 0;
        if (getenv("SA_VERBOSE"))
            fprintf(stderr, "  device-resident alignment loop: %zu pairs, device call %.3f s, host replay + files %.3f s, %zu pairs left to batched rounds\n",
                    n_pairs, t_r1 - t_r0, wall_now() - t_r1, n_fallback);
        if (n_bad) {
            fprintf(stderr, "SegAligner(b200): internal error: device alignment loop and host replay disagree on %d pairs\n", n_bad);
            throw JobExit{2, true};
        }
    }

    Engine::AlignOut out;  // reused across rounds
    std::vector<size_t> live;
    std::vector<sa_problem> pr;
    for (;;) {
        live.clear();
        for (size_t k = 0; k < st.size(); k++) {
            PairState &p = st[k];
            if (!p.live) continue;
            if (p.curr_best >= p.exp_thresh && p.count < max_alns) live.push_back(k);  // main.cpp:268
            else p.live = false;
        }
        if (live.empty()) break;
        if (arena.size() >= (1ull << 31)) {
            fprintf(stderr, "SegAligner(b200): skip bitmaps of this run exceed 2^31 words\n");
            throw JobExit{2, true};
        }
        pr.resize(live.size());
        for (size_t n = 0; n < live.size(); n++) {
            const PairState &p = st[live[n]];
            pr[n] = sa_problem{p.c, p.s, (int32_t)p.used_off, 0};
        }
        const double t_a0 = wall_now();
        eng.align(mode, pr, arena, out);
        const double t_a1 = wall_now();
        struct RoundLog {
            double t0, t1;
            size_t n;
            ~RoundLog() {
                if (getenv("SA_VERBOSE"))
                    fprintf(stderr, "  alignment round: %zu pairs, align call %.3f s, host post-processing %.3f s\n", n, t1 - t0,
                            wall_now() - t1);
            }
        } round_log{t_a0, t_a1, live.size()};
        // Post-processing of the round, in parallel over contiguous ranges of pairs.  Pairs are independent except for
        // (a) new private bitmaps, which each worker appends to its own arena piece (spliced into the arena afterwards)
        // and (b) the labels of WRITTEN alignments, which each worker records and the main thread ORs in afterwards.
        const size_t n_live = live.size();
        const int n_workers = (int)std::max<size_t>(1, std::min<size_t>(std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u),
                                                                        n_live / 50000 + 1));
        struct Worker {
            std::vector<uint32_t> piece;           // private bitmaps created this round
            std::vector<size_t> fresh;             // pairs (index into st) whose used_off is relative to `piece`
            std::vector<std::pair<int, int>> disc;  // (contig, label) of written alignments
            std::vector<AlnEntry> aln;
        };
        std::vector<Worker> workers((size_t)n_workers);
        auto work = [&](int wi) {
            Worker &W = workers[(size_t)wi];
            const size_t lo = n_live * (size_t)wi / (size_t)n_workers, hi = n_live * ((size_t)wi + 1) / (size_t)n_workers;
            std::vector<AlnEntry> &aln = W.aln;
            for (size_t n = lo; n < hi; n++) {
                PairState &p = st[live[n]];
                const Engine::AlignOut::View v = out.get(n);
                const int len = v.res->path_len;
                if (len <= 0) {  // no backtrack start (cannot happen with >= 1 label each); treat as exhausted
                    p.live = false;
                    continue;
                }
                const int32_t *pj = v.pj;
                p.exp_thresh = partial_thresh(p, len, pj);
                p.curr_best = v.res->score;
                if (!(p.curr_best > p.exp_thresh)) {
                    // equality would spin forever in the reference (same DP, same result); we stop instead
                    if (p.curr_best == p.exp_thresh) p.live = false;
                    continue;
                }
                p.count += 1;
                uint32_t *bits;
                if (!p.own) {  // first accepted alignment of the pair: private copy of the seed (or an empty bitmap)
                    const size_t w = bm_words(contigs.labels(p.c));
                    const size_t at = W.piece.size();
                    W.piece.resize(at + w, 0u);
                    if (p.used_off >= 0) std::copy(arena.begin() + p.used_off, arena.begin() + p.used_off + (long)w, W.piece.begin() + (long)at);
                    p.used_off = (int64_t)at;  // relative to the piece until the splice below
                    p.own = true;
                    W.fresh.push_back(live[n]);
                    bits = W.piece.data() + at;
                } else {
                    bits = arena.data() + p.used_off;  // the arena itself is not resized while the workers run
                }
                for (int t = 0; t < len; t++) bits[(size_t)pj[t] >> 5] |= 1u << (pj[t] & 31);  // main.cpp:298-300
                emit(p, len, pj, v.pq, v.ps, aln, W.disc);
            }
        };
        if (n_workers == 1) {
            work(0);
        } else {
            std::vector<std::thread> th;
            for (int wi = 0; wi < n_workers; wi++) th.emplace_back(work, wi);
            for (auto &t : th) t.join();
        }
        for (Worker &W : workers) {  // splice the new private bitmaps into the arena, record the written labels
            const int64_t base = (int64_t)arena.size();
            arena.insert(arena.end(), W.piece.begin(), W.piece.end());
            for (size_t k : W.fresh) st[k].used_off += base;
            for (auto &e : W.disc) bm_set(discovered[(size_t)e.first], e.second);
        }
    }
    return discovered;
}

double wall_now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
struct PhaseTimer {  // SA_VERBOSE=1: per-phase wall times on stderr
    bool on = getenv("SA_VERBOSE") != nullptr;
    double t = wall_now();
    void lap(const char *what) {
        if (!on) return;
        const double n = wall_now();
        fprintf(stderr, "SegAligner(b200): %-28s %8.3f s\n", what, n - t);
        t = n;
    }
};

}  // namespace

// One SegAligner run (the body of the reference's main, main.cpp:453-656).  `eng` may already hold open contexts
// (resident server); `keep_open` leaves them open on return.
static int run_job(int argc, char *argv[], Engine &eng, bool keep_open) {
    if (argc < 3) {  // the reference tests argc < 2 and then dereferences argv[2] regardless (main.cpp:454-468)
        printf("wrong number of arguments\n");
        fflush(stdout);
        throw JobExit{1, false};
    }
    const clock_t begin = clock();
    const double w0 = wall_now();
    const std::string ref_cmap_file = argv[1];
    const std::string contig_cmap_file = argv[2];
    Options o;
    parse_args(argc, argv, o);
    if (!o.limit_lookback) {
        o.lookback = 9999999;
        printf("Alignment banding OFF. \n");
    }
    if (o.inst_gen == 2) o.max_seg_contig_alns *= 2;  // main.cpp:479-481

    PhaseTimer pt;
    // A run that was told to use one GPU hides the others from the CUDA driver before its first CUDA call: driver
    // initialisation enumerates every visible device, which costs seconds on an 8-GPU box (measured: 5-6 s to open
    // one context there vs 0.4-2 s on a single-GPU box).  Only when nothing is open yet (not in the resident server).
    {
        int want = o.n_gpus;
        if (want <= 0)
            if (const char *e = getenv("SA_NGPUS")) want = atoi(e);
        if (want == 1 && eng.ctx.empty() && !keep_open) {
            const char *vis = getenv("CUDA_VISIBLE_DEVICES");
            std::string first = "0";
            if (vis && *vis) {
                first = vis;
                const size_t comma = first.find(',');
                if (comma != std::string::npos) first.resize(comma);
            }
            setenv("CUDA_VISIBLE_DEVICES", first.c_str(), 1);
        }
    }
    // CUDA start-up (driver init + context, ~0.3 s) overlaps with CMAP parsing: open device 0 right away
    JobExit open_err{0, false};
    std::thread gpu_open([&eng, &open_err]() {
        try {
            eng.open(1);
        } catch (const JobExit &e) {
            open_err = e;
        }
    });
    struct Joiner {  // the opener thread must be joined on every exit path (exceptions included)
        std::thread &t;
        ~Joiner() {
            if (t.joinable()) t.join();
        }
    } joiner{gpu_open};
    MapSet segs_raw, segs, contigs;
    parse_cmap(ref_cmap_file, segs_raw);
    if (!o.fitting_aln) segs = make_reverse_cmap(segs_raw, o.min_map_len);
    else segs = segs_raw;
    parse_cmap(contig_cmap_file, contigs);
    non_collapse_probs(segs, o.inst_gen);
    non_collapse_probs(contigs, o.inst_gen);
    pt.lap("parse + prepare maps");

    // contig set (get_contig_set, main.cpp:437-447)
    std::set<int> contig_set;
    if (!o.contig_list_file.empty()) {
        for (int id : parse_contig_list(o.contig_list_file)) contig_set.insert(id);
    } else {
        for (int id : contigs.ids) contig_set.insert(id);
    }
    printf("Running SegAligner on %d segments and %d contigs.\n", segs.n_maps(), contigs.n_maps());
    fflush(stdout);
    if (segs.n_maps() == 0 || contigs.n_maps() == 0) {
        // nothing to align; the reference still writes the (empty) score files unless -fitting
        if (!o.fitting_aln) {  // same progress lines as the reference's (empty) phases (main.cpp:528-649)
            printf("Computing scoring thresholds.\n");
            write_all_scores({}, o.sample_prefix);
            printf("Writing scoring distribution.\n");
            write_score_thresholds({}, o.sample_prefix + o.detection_label);
            printf("Completed generating scores.\n");
            printf("elapsed seconds for scoring %f\n", double(clock() - begin) / CLOCKS_PER_SEC);
            printf("Performing alignments\nFinished standard alignment. \n");
            if (o.do_tip) printf("Doing tip alignment.\nPerforming alignments\nFinished tip alignments.\n\n");
        }
        printf("elapsed seconds for run %f\n", double(clock() - begin) / CLOCKS_PER_SEC);
        fflush(stdout);
        gpu_open.join();
        if (open_err.code) throw open_err;
        if (!keep_open) eng.close();
        return 0;
    }

    // GPUs: all visible ones for big runs, one otherwise (context creation costs more than a small run)
    double total_cells = 0;
    {
        double sl = 0, cl = 0;
        for (int m = 0; m < segs.n_maps(); m++) sl += segs.labels(m);
        for (int m = 0; m < contigs.n_maps(); m++) cl += contigs.labels(m);
        total_cells = sl * cl;
    }
    int n_gpus = o.n_gpus;
    if (const char *e = getenv("SA_NGPUS"))
        if (n_gpus <= 0) n_gpus = atoi(e);
    if (n_gpus <= 0) n_gpus = total_cells > 2e10 ? sa_device_count() : 1;
    gpu_open.join();
    if (open_err.code) throw open_err;
    if (n_gpus > 1) eng.open_more(n_gpus);
    pt.lap("open GPU context(s)");
    eng.set_maps(contigs, segs);
    pt.lap("upload maps");

    // ---------------- fitting mode (main.cpp:512-526, run_SA_fitting :169-208) --------------------------------
    if (o.fitting_aln) {
        std::vector<sa_problem> pr;
        for (int id : contig_set) {
            const int c = contigs.index_of(id);
            for (int s = 0; s < segs.n_maps(); s++) {
                if (c < 0) continue;  // the reference would create an empty map here and crash; skip
                pr.push_back(sa_problem{c, s, -1, 0});
            }
        }
        sa_mode mode{SA_INIT_FITTING, SA_START_END, 1, o.lookback, o.swap_b_x ? 1 : 0};
        Engine::AlignOut out;
        eng.align(mode, pr, std::vector<uint32_t>(), out);
        std::vector<AlnEntry> aln;
        for (size_t k = 0; k < pr.size(); k++) {
            const int x = segs.ids[(size_t)pr[k].seg];
            const int contig_id = contigs.ids[(size_t)pr[k].contig];
            const std::string outname = o.sample_prefix + "_" + std::to_string(contig_id) + "_" +
                                        std::to_string(x < 0 ? -x : x) + "_fitting_aln.txt";
            const Engine::AlignOut::View v = out.get(k);
            aln.resize((size_t)std::max(v.res->path_len, 0));
            for (size_t t = 0; t < aln.size(); t++) aln[t] = AlnEntry{v.pj[t], v.pq[t], v.ps[t]};
            print_alignment(outname, aln, contig_id, x, segs.size(pr[k].seg), v.res->score);
        }
        printf("elapsed seconds for run %f\n", double(clock() - begin) / CLOCKS_PER_SEC);
        fflush(stdout);
        if (!keep_open) eng.close();
        return 0;
    }

    // ---------------- phase 1: scoring (main.cpp:528-553, run_SA_score :112-167) ------------------------------
    printf("Computing scoring thresholds.\n");
    fflush(stdout);
    const int nt = std::max(o.n_threads, 1);
    std::vector<std::vector<int>> chunk((size_t)nt);  // contig indices per reference thread, ascending
    {
        int count = 0;
        for (int id : contig_set) {
            const int c = contigs.index_of(id);
            if (c >= 0) chunk[(size_t)(count % nt)].push_back(c);
            count++;
        }
    }
    // problems in the reference's row order: thread, then segment, then contig
    std::vector<sa_problem> pr;
    for (int t = 0; t < nt; t++)
        for (int s = 0; s < segs.n_maps(); s++)
            for (int c : chunk[(size_t)t]) pr.push_back(sa_problem{c, s, -1, 0});
    std::vector<ScoreRow> rows;
    {
        sa_mode mode;
        mode.init_mode = SA_INIT_SEMIGLOBAL;  // "handles both cases" (main.cpp:141)
        mode.start_mode = o.local_aln ? SA_START_LOCAL : SA_START_SEMIGLOBAL;  // tip_aln is false here (:146-152)
        mode.fitting = 0;
        mode.lookback = o.lookback;
        mode.swap_b_x = o.swap_b_x ? 1 : 0;
        if (o.detection) {
            int total_labels = 0;
            for (int c = 0; c < contigs.n_maps(); c++) total_labels += contigs.size(c);  // main.cpp:118-121
            std::vector<int32_t> want(pr.size());
            for (size_t k = 0; k < pr.size(); k++)
                want[k] = int(roundf(o.n_detect_scores * float(contigs.size(pr[k].contig)) / float(total_labels))) + 1;
            std::vector<std::vector<float>> sc;
            eng.detect(mode, pr, want, sc);
            for (size_t k = 0; k < pr.size(); k++)
                for (float v : sc[k]) rows.push_back(ScoreRow{segs.ids[(size_t)pr[k].seg], contigs.ids[(size_t)pr[k].contig], v});
        } else {
            std::vector<sa_result> res;
            eng.score(mode, pr, res);
            rows.reserve(pr.size());
            for (size_t k = 0; k < pr.size(); k++)
                rows.push_back(ScoreRow{segs.ids[(size_t)pr[k].seg], contigs.ids[(size_t)pr[k].contig], res[k].score});
        }
    }
    pt.lap("phase 1 scoring (GPU)");
    write_all_scores(rows, o.sample_prefix);
    pt.lap("write all_scores");

    // ---------------- phase 2: thresholds (main.cpp:555-557) ---------------------------------------------------
    std::vector<std::vector<std::pair<int, float>>> thr2;
    compute_score_thresholds_multi(rows, segs, contigs, {(double)o.p_val, (double)o.p_val_tip}, o.n_detect_scores, thr2);
    const auto &score_thresholds = thr2[0];
    const auto &tip_thresholds = thr2[1];
    printf("Writing scoring distribution.\n");
    write_score_thresholds(score_thresholds, o.sample_prefix + o.detection_label);
    printf("Completed generating scores.\n");
    printf("elapsed seconds for scoring %f\n", double(clock() - begin) / CLOCKS_PER_SEC);
    fflush(stdout);

    pt.lap("thresholds");
    // ---------------- phase 3: alignment of the pairs that pass (main.cpp:563-604) ----------------------------
    std::vector<std::pair<int, int>> pairs_list;  // (seg idx, contig idx), in row order, duplicates kept
    for (auto &r : rows) {
        const int s = segs.index_of(r.seg);
        if (r.score >= score_thresholds[(size_t)s].second) pairs_list.emplace_back(s, contigs.index_of(r.contig));
    }
    std::vector<std::pair<int, int>> uniq = pairs_list;
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    printf("Performing alignments\n");
    fflush(stdout);
    std::vector<Bitmap> contig_used = run_alignment_pass(eng, o, segs, contigs, uniq, score_thresholds, nullptr, false);
    printf("Finished standard alignment. \n");
    fflush(stdout);
    pt.lap("phase 3 alignment");

    // ---------------- phase 4: tip alignment (main.cpp:606-649) ------------------------------------------------
    if (o.do_tip) {
        printf("Doing tip alignment.\n");
        std::set<int> tip_contigs;
        for (auto &e : pairs_list)
            if (bm_any(contig_used[(size_t)e.second])) tip_contigs.insert(e.second);
        std::vector<std::pair<int, int>> tip_pairs;
        for (int s = 0; s < segs.n_maps(); s++)
            for (int c : tip_contigs) tip_pairs.emplace_back(s, c);  // every segment (+ and -) x every such contig
        printf("Performing alignments\n");
        fflush(stdout);
        run_alignment_pass(eng, o, segs, contigs, tip_pairs, tip_thresholds, &contig_used, true);
        printf("Finished tip alignments.\n\n");
        pt.lap("phase 4 tip alignment");
    }
    printf("elapsed seconds for run %f\n", double(clock() - begin) / CLOCKS_PER_SEC);
    if (getenv("SA_VERBOSE")) {
        uint64_t h2d = 0, d2h = 0;
        for (sa_ctx *c : eng.ctx) {
            uint64_t a = 0, b = 0;
            sa_get_traffic(c, &a, &b);
            h2d += a;
            d2h += b;
        }
        fprintf(stderr, "SegAligner(b200): copied H2D %llu bytes, D2H %llu bytes\n", (unsigned long long)h2d, (unsigned long long)d2h);
        fprintf(stderr, "SegAligner(b200): wall %.3f s on %d GPU(s)\n", wall_now() - w0, (int)eng.ctx.size());
    }
    fflush(stdout);
    fflush(stderr);
    if (!keep_open) eng.close();  // orderly teardown: abandoning the context only moves the cost to the next start
    pt.lap("close");
    return 0;
}

// ---- resident server ----------------------------------------------------------------------------------------------
// OMPathFinder.py:407-428 execs one `SegAligner -fitting` process per candidate path (a 50-cell problem, 0.2 ms in
// the reference); creating a CUDA context costs 0.3-6 s per process.  The first such invocation therefore leaves a
// background server holding the GPU context on a Unix socket; later invocations forward argv + cwd to it and print the
// stdout it returns.  Same executable, same files, same exit codes; no CPU fallback anywhere.
//   -fitting runs           use the server by default (idle timeout 30 s); SA_SERVER=0 turns that off
//   SA_SERVER=1             every run goes through the server (idle timeout 300 s); SA_SERVER_IDLE=<s> overrides both
// The socket lives in a directory only this user can enter ($XDG_RUNTIME_DIR/segaligner_b200, else
// /tmp/segaligner_b200_<uid>, mode 0700, owner and mode verified with lstat), is bound under umask 077, and both ends
// check SO_PEERCRED: a peer with another uid is never trusted -- the client then runs the job in-process instead.
#include <fcntl.h>
#include <poll.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <sys/wait.h>

namespace {

// a directory that exists (created 0700 if missing), is a real directory (no symlink), is ours and is closed to others
bool private_dir_ok(const std::string &dir) {
    if (mkdir(dir.c_str(), 0700) != 0 && errno != EEXIST) return false;
    struct stat sb;
    if (lstat(dir.c_str(), &sb) != 0) return false;
    return S_ISDIR(sb.st_mode) && sb.st_uid == getuid() && (sb.st_mode & 077) == 0;
}

// "" = no trustworthy place for the socket (the caller runs in-process)
std::string server_socket_path() {
    if (const char *e = getenv("SA_SERVER_SOCKET")) return e;  // explicit path: the caller vouches for the directory
    std::string dir;
    if (const char *x = getenv("XDG_RUNTIME_DIR")) {
        struct stat sb;
        if (*x && lstat(x, &sb) == 0 && S_ISDIR(sb.st_mode) && sb.st_uid == getuid() && (sb.st_mode & 077) == 0)
            dir = std::string(x) + "/segaligner_b200";
    }
    if (dir.empty()) dir = "/tmp/segaligner_b200_" + std::to_string((long)getuid());
    if (!private_dir_ok(dir)) return "";
    return dir + "/server.sock";
}

bool peer_is_me(int fd) {
    struct ucred cr;
    socklen_t len = sizeof cr;
    if (getsockopt(fd, SOL_SOCKET, SO_PEERCRED, &cr, &len) != 0 || len != sizeof cr) return false;
    return cr.uid == getuid();
}

bool write_all(int fd, const void *buf, size_t n) {
    const char *p = (const char *)buf;
    while (n) {
        const ssize_t w = write(fd, p, n);
        if (w <= 0) return false;
        p += w;
        n -= (size_t)w;
    }
    return true;
}
bool read_all(int fd, void *buf, size_t n) {
    char *p = (char *)buf;
    while (n) {
        const ssize_t r = read(fd, p, n);
        if (r <= 0) return false;
        p += r;
        n -= (size_t)r;
    }
    return true;
}
bool recv_str(int fd, std::string &s) {
    uint32_t n = 0;
    if (!read_all(fd, &n, 4) || n > (64u << 20)) return false;
    s.resize(n);
    return n == 0 || read_all(fd, &s[0], n);
}

int connect_server(const std::string &path) {
    const int fd = socket(AF_UNIX, SOCK_STREAM, 0);
    if (fd < 0) return -1;
    sockaddr_un a{};
    a.sun_family = AF_UNIX;
    snprintf(a.sun_path, sizeof a.sun_path, "%s", path.c_str());
    if (connect(fd, (sockaddr *)&a, sizeof a) != 0) {
        close(fd);
        return -1;
    }
    if (!peer_is_me(fd)) {  // somebody else's process answers on this path: never hand it our argv or trust its reply
        close(fd);
        return -2;
    }
    return fd;
}

[[noreturn]] void serve(const std::string &path, int idle_s) {
    const int ls = socket(AF_UNIX, SOCK_STREAM, 0);
    sockaddr_un a{};
    a.sun_family = AF_UNIX;
    snprintf(a.sun_path, sizeof a.sun_path, "%s", path.c_str());
    umask(077);  // the socket is created 0600: no window between bind() and a later chmod()
    if (ls < 0 || bind(ls, (sockaddr *)&a, sizeof a) != 0 || listen(ls, 64) != 0) _exit(0);  // somebody else serves
    if (const char *e = getenv("SA_SERVER_IDLE")) idle_s = std::max(1, atoi(e));
    // The server drives one GPU unless its environment asks for more (SA_NGPUS > 1): hiding the other devices before
    // the first CUDA call saves seconds of driver initialisation on a multi-GPU box (see run_job).
    {
        const char *ng = getenv("SA_NGPUS");
        if (!ng || atoi(ng) <= 1) {
            const char *vis = getenv("CUDA_VISIBLE_DEVICES");
            std::string first = "0";
            if (vis && *vis) {
                first = vis;
                const size_t comma = first.find(',');
                if (comma != std::string::npos) first.resize(comma);
            }
            setenv("CUDA_VISIBLE_DEVICES", first.c_str(), 1);
        }
    }
    Engine eng;
    bool fatal = false;
    while (!fatal) {
        pollfd pf{ls, POLLIN, 0};
        if (poll(&pf, 1, idle_s * 1000) <= 0) break;  // idle: give the GPU back
        const int c = accept(ls, nullptr, nullptr);
        if (c < 0) continue;
        if (!peer_is_me(c)) {
            close(c);
            continue;
        }
        uint32_t n_args = 0;
        std::vector<std::string> args;
        std::string cwd;
        bool ok = read_all(c, &n_args, 4) && n_args < 4096;
        for (uint32_t k = 0; ok && k < n_args; k++) {
            std::string s;
            ok = recv_str(c, s);
            args.push_back(s);
        }
        ok = ok && recv_str(c, cwd);
        if (!ok) {
            close(c);
            continue;
        }
        if (args.size() == 2 && args[1] == "--server-stop") {
            const uint32_t code = 0;
            const uint64_t zero = 0;
            write_all(c, &code, 4);
            write_all(c, &zero, 8);
            close(c);
            break;
        }
        // run the job with stdout captured into an anonymous memory file (no file-system round trip)
        const int tf = memfd_create("sa_b200_stdout", MFD_CLOEXEC);
        if (tf < 0) {
            close(c);
            continue;
        }
        fflush(stdout);
        const int saved = dup(1);
        dup2(tf, 1);
        if (chdir(cwd.c_str()) != 0) { /* paths relative to an unreachable cwd will simply fail to open */ }
        std::vector<char *> av;
        for (auto &s : args) av.push_back(&s[0]);
        av.push_back(nullptr);
        uint32_t code = 0;
        try {
            code = (uint32_t)run_job((int)args.size(), av.data(), eng, true);
        } catch (const JobExit &e) {
            code = (uint32_t)e.code;
            fatal = e.fatal;
        }
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
        const off_t len = lseek(tf, 0, SEEK_END);
        lseek(tf, 0, SEEK_SET);
        std::string out((size_t)std::max<off_t>(len, 0), '\0');
        if (len > 0) read_all(tf, &out[0], (size_t)len);
        close(tf);
        const uint64_t olen = out.size();
        write_all(c, &code, 4);
        write_all(c, &olen, 8);
        write_all(c, out.data(), out.size());
        close(c);
    }
    close(ls);
    unlink(path.c_str());
    eng.close();
    _exit(0);
}

// returns -1 when the server route is unavailable or not trustworthy (caller then runs the job in-process)
int run_via_server(int argc, char *argv[], int idle_s) {
    const std::string path = server_socket_path();
    if (path.empty() || path.size() >= sizeof(sockaddr_un{}.sun_path)) return -1;
    int fd = connect_server(path);
    if (fd == -2) return -1;
    if (fd < 0 && argc == 2 && strcmp(argv[1], "--server-stop") == 0) return 0;  // nothing to stop
    if (fd < 0) {
        // no server yet (or a stale socket): start one.  fork happens before any CUDA call in this process.
        unlink(path.c_str());
        const pid_t pid = fork();
        if (pid == 0) {
            setsid();
            signal(SIGHUP, SIG_IGN);
            signal(SIGPIPE, SIG_IGN);
            const int dn = open("/dev/null", O_RDWR);
            if (dn >= 0) {
                dup2(dn, 0);
                dup2(dn, 1);
                if (!getenv("SA_VERBOSE")) dup2(dn, 2);
            }
            for (int f = 3; f < 256; f++) close(f);
            serve(path, idle_s);
        }
        if (pid < 0) return -1;
        for (int tries = 0; tries < 3000 && fd < 0; tries++) {  // the socket appears before CUDA is initialised
            usleep(1000);
            fd = connect_server(path);
            if (fd == -2) return -1;
        }
        if (fd < 0) return -1;
    }
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) cwd[0] = 0;
    // one write: argc, argv strings and cwd, each length-prefixed
    std::string msg;
    auto put = [&msg](const void *p, size_t n) { msg.append((const char *)p, n); };
    const uint32_t n_args = (uint32_t)argc;
    put(&n_args, 4);
    for (int k = 0; k < argc; k++) {
        const uint32_t n = (uint32_t)strlen(argv[k]);
        put(&n, 4);
        put(argv[k], n);
    }
    const uint32_t nc = (uint32_t)strlen(cwd);
    put(&nc, 4);
    put(cwd, nc);
    bool ok = write_all(fd, msg.data(), msg.size());
    uint32_t code = 0;
    uint64_t olen = 0;
    ok = ok && read_all(fd, &code, 4) && read_all(fd, &olen, 8) && olen < (1ull << 32);
    if (!ok) {
        close(fd);
        fprintf(stderr, "SegAligner(b200): the resident server went away while running this job\n");
        return 134;
    }
    std::string out((size_t)olen, '\0');
    if (olen) read_all(fd, &out[0], (size_t)olen);
    close(fd);
    fwrite(out.data(), 1, out.size(), stdout);
    fflush(stdout);
    if (code == 2)
        fprintf(stderr, "SegAligner(b200): the job failed inside the resident server (exit code 2); rerun with SA_SERVER=0 "
                        "to see its diagnostics\n");
    return (int)code;
}

}  // namespace

int main(int argc, char *argv[]) {
    signal(SIGPIPE, SIG_IGN);
    // the resident server: always with SA_SERVER=1, never with SA_SERVER=0, by default for -fitting runs (the
    // OMPathFinder pattern: thousands of tiny one-shot processes) with a short idle timeout
    const char *sv = getenv("SA_SERVER");
    const bool forced = sv && atoi(sv) != 0, forbidden = sv && atoi(sv) == 0;
    bool fitting = false;
    for (int i = 3; i < argc; i++) fitting = fitting || strcmp(argv[i], "-fitting") == 0;
    if (forced || (fitting && !forbidden) || (argc == 2 && strcmp(argv[1], "--server-stop") == 0)) {
        const int rc = run_via_server(argc, argv, forced ? 300 : 30);
        if (rc >= 0) return rc;
    }
    Engine eng;
    try {
        return run_job(argc, argv, eng, false);
    } catch (const JobExit &e) {
        fflush(stdout);
        return e.code;
    }
}
