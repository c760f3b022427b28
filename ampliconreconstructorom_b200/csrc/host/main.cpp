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
#include <atomic>
#include <thread>
#include <unordered_map>
#include <vector>
#include <fcntl.h>
#include <sched.h>
#include <sys/wait.h>
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

// ---- one process per GPU -------------------------------------------------------------------------------------------
// A multi-GPU run is SPMD: main() forks one process per GPU BEFORE the first CUDA call, each process sees exactly one device
// (CUDA_VISIBLE_DEVICES), parses the inputs itself and runs the same run_job; they differ only in which shard of every
// problem list they compute (problems sharded by index, no device collective -- the problems are independent).  Driver
// initialisation then costs what it costs for ONE device, in parallel, instead of enumerating every GPU of the box in one
// process (measured: 5-6 s on an 8-GPU box), and the host work of the alignment passes (replay, filters, ~10^4 output files)
// is spread over the processes too.  The three exchange points of SURVEY.md 8(e) go through a MAP_SHARED region created
// before the fork: (1) all phase-1 scores, (2) the used contig labels after the standard alignment pass, (3) the end.
struct SpmdCtl {  // first page of the shared region
    std::atomic<int> arrived, generation, error;
    std::atomic<unsigned long long> h2d, d2h;
    std::atomic<int> init_turn;        // CUDA start-up goes rank by rank (see Spmd::wait_init_turn)
    std::atomic<long long> next_unit;  // phase 1: cursor into the shared problem order (see score_dynamic)
};
struct Spmd {
    int rank = 0, world = 1;  // world starts as the number of processes forked and may shrink to 1 for small runs
    SpmdCtl *ctl = nullptr;
    char *region = nullptr;
    size_t region_bytes = 0, top = 4096;
    std::vector<pid_t> kids;  // rank 0 only

    // Sense-reversing barrier over the first `world` ranks.  A rank that fails sets ctl->error, which releases everybody
    // else with exit code 2; rank 0 also notices children that died without saying so.
    void barrier() {
        if (world <= 1) return;
        const int gen = ctl->generation.load();
        if (ctl->arrived.fetch_add(1) + 1 == world) {
            ctl->arrived.store(0);
            ctl->generation.fetch_add(1);
            return;
        }
        for (unsigned spin = 0; ctl->generation.load() == gen; spin++) {
            if (ctl->error.load()) throw JobExit{2, true};
            if (rank == 0 && (spin & 1023) == 1023) {
                for (pid_t k : kids) {
                    int st = 0;
                    if (waitpid(k, &st, WNOHANG) == k && !(WIFEXITED(st) && WEXITSTATUS(st) == 0)) {
                        ctl->error.store(1);
                        fprintf(stderr, "SegAligner(b200): a GPU worker process died\n");
                        throw JobExit{2, true};
                    }
                }
            }
            if (spin < 2000) sched_yield();
            else usleep(100);
        }
        if (ctl->error.load()) throw JobExit{2, true};
    }
    // CUDA start-up (driver initialisation + context creation) is serialised by the driver across the processes of a box:
    // measured on a 4 x B200 box (profiles/r2_cuda_startup_probe_4gpu.txt) one process alone is ready after 0.5 s, four
    // processes starting together are ALL ready only after 2.5-3.2 s, while four processes taking turns are ready after 0.5,
    // 1.3, 2.1 and 2.9 s.  So the ranks take turns, rank 0 first, and phase 1 hands out work dynamically: the ranks that are
    // up early start scoring while the later ones are still initialising.
    void wait_init_turn() {
        if (world <= 1 || !ctl) return;
        for (unsigned spin = 0; ctl->init_turn.load() < rank; spin++) {
            if (ctl->error.load()) throw JobExit{2, true};
            usleep(spin < 50 ? 200 : 2000);
        }
    }
    void pass_init_turn() {
        if (!ctl) return;
        int cur = ctl->init_turn.load();
        while (cur <= rank && !ctl->init_turn.compare_exchange_weak(cur, rank + 1)) {
        }
    }
    // Bump allocation inside the shared region; every rank calls it with the same sizes in the same order, so they all
    // get the same offsets without talking to each other.
    template <class T>
    T *alloc(size_t n) {
        top = (top + 63) & ~(size_t)63;
        const size_t at = top;
        top += n * sizeof(T);
        if (world > 1 && top > region_bytes) {
            fprintf(stderr, "SegAligner(b200): the inter-process exchange region (%zu bytes) is too small\n", region_bytes);
            throw JobExit{2, true};
        }
        return reinterpret_cast<T *>(region + at);
    }
};

// shard[g] = indices of the problems rank g handles.  Problems are dealt round-robin WITHIN size classes (1/4-octave of
// cells), largest classes first, so every GPU gets the same mix of sizes in O(n) (the library orders each shard
// largest-first itself); a full sort of 10^7 problems would cost more than their scoring.  Deterministic: every rank
// computes the same partition.
std::vector<int64_t> my_shard(const std::vector<sa_problem> &pr, const MapSet &contigs, const MapSet &segs, int world, int rank) {
    std::vector<int64_t> mine;
    if (world <= 1) {
        mine.resize(pr.size());
        for (size_t k = 0; k < pr.size(); k++) mine[k] = (int64_t)k;
        return mine;
    }
    const int NB = 4 * 64;
    std::vector<uint8_t> cls(pr.size());
    std::vector<int64_t> cnt((size_t)NB + 1, 0);
    for (size_t k = 0; k < pr.size(); k++) {
        const double cells = (double)contigs.labels(pr[k].contig) * (double)segs.labels(pr[k].seg);
        int b = cells > 1.0 ? (int)(std::log2(cells) * 4.0) : 0;
        b = NB - 1 - std::min(b, NB - 1);  // descending size
        cls[k] = (uint8_t)b;
        cnt[(size_t)b + 1]++;
    }
    for (int b = 0; b < NB; b++) cnt[(size_t)b + 1] += cnt[(size_t)b];
    mine.reserve(pr.size() / (size_t)world + 1);
    for (size_t k = 0; k < pr.size(); k++) {
        const int64_t r = cnt[cls[k]]++;  // position in the size-class order
        if (r % world == rank) mine.push_back((int64_t)k);
    }
    return mine;
}

// The same size-class order as a list: order = problem indices, largest classes first (stable inside a class);
// cum[k] = cells of order[0 .. k).  Every rank computes the same arrays.
void size_class_order(const std::vector<sa_problem> &pr, const MapSet &contigs, const MapSet &segs, std::vector<int64_t> &order,
                      std::vector<double> &cum) {
    const int NB = 4 * 64;
    std::vector<uint8_t> cls(pr.size());
    std::vector<int64_t> cnt((size_t)NB + 1, 0);
    std::vector<double> cells(pr.size());
    for (size_t k = 0; k < pr.size(); k++) {
        cells[k] = (double)contigs.labels(pr[k].contig) * (double)segs.labels(pr[k].seg);
        int b = cells[k] > 1.0 ? (int)(std::log2(cells[k]) * 4.0) : 0;
        b = NB - 1 - std::min(b, NB - 1);
        cls[k] = (uint8_t)b;
        cnt[(size_t)b + 1]++;
    }
    for (int b = 0; b < NB; b++) cnt[(size_t)b + 1] += cnt[(size_t)b];
    order.resize(pr.size());
    for (size_t k = 0; k < pr.size(); k++) order[(size_t)cnt[cls[k]]++] = (int64_t)k;
    cum.assign(pr.size() + 1, 0.0);
    for (size_t k = 0; k < pr.size(); k++) cum[k + 1] = cum[k] + cells[(size_t)order[k]];
}

// ---- this process's GPU: one context behind the C ABI ----------------------------------------------------------------
struct Engine {
    sa_ctx *ctx = nullptr;

    void die(const char *what, int rc) {
        fprintf(stderr, "SegAligner(b200): %s failed (%d): %s\n", what, rc, ctx ? sa_last_error(ctx) : "");
        throw JobExit{2, true};
    }
    void open() {
        if (ctx) return;  // resident server: the context stays open between jobs
        if (sa_device_count() <= 0) {
            fprintf(stderr, "SegAligner(b200): no CUDA device visible; this build has no CPU fallback\n");
            throw JobExit{2, true};
        }
        const int rc = sa_ctx_create(0, &ctx);
        if (rc) {
            ctx = nullptr;
            die("sa_ctx_create (need an sm_100 GPU)", rc);
        }
    }
    void close() {
        if (ctx) sa_ctx_destroy(ctx);
        ctx = nullptr;
    }
    void set_maps(const MapSet &c, const MapSet &s) {
        int rc = sa_set_maps(ctx, SA_SIDE_CONTIG, c.n_maps(), c.off.data(), c.pos.data(), c.prob.data());
        if (rc) die("sa_set_maps(contigs)", rc);
        rc = sa_set_maps(ctx, SA_SIDE_SEG, s.n_maps(), s.off.data(), s.pos.data(), s.prob.data());
        if (rc) die("sa_set_maps(segments)", rc);
    }
    void score(const sa_mode &mode, const std::vector<sa_problem> &pr, std::vector<sa_result> &res) {
        res.assign(pr.size(), sa_result{0, -1, -1, 0});
        if (pr.empty()) return;
        const int rc = sa_score_batch(ctx, &mode, pr.data(), (int64_t)pr.size(), res.data());
        if (rc) die("sa_score_batch", rc);
    }
    // detection scoring: sc[off[k] .. off[k] + got[k]) = the scores multi_backtrack_scores harvests for problem k, in order
    void detect(const sa_mode &mode, const std::vector<sa_problem> &pr, const std::vector<int32_t> &want, std::vector<int64_t> &off,
                std::vector<float> &sc, std::vector<int32_t> &got) {
        const size_t m = pr.size();
        off.assign(m + 1, 0);
        for (size_t k = 0; k < m; k++) off[k + 1] = off[k] + want[k];
        sc.assign((size_t)std::max<int64_t>(off[m], 1), 0.0f);
        got.assign(m, 0);
        if (m == 0) return;
        const int rc = sa_detect_batch(ctx, &mode, pr.data(), (int64_t)m, want.data(), off.data(), sc.data(), got.data());
        if (rc) die("sa_detect_batch", rc);
    }
    // alignment: out.get(k) = result and traceback (end->start) of problem k, read in place from the context's pinned
    // buffers (valid until the next align call).
    struct AlignOut {
        struct View {
            const sa_result *res;
            const int32_t *pj, *pq;
            const float *ps;
        };
        sa_packed_paths pk{};
        View get(size_t k) const { return View{pk.results + k, pk.path_j + pk.start[k], pk.path_q + pk.start[k], pk.path_s + pk.start[k]}; }
    };
    // pr[k].used_slot = WORD offset of the problem's skip bitmap inside `arena` (or -1): the C ABI is used with one
    // word per slot, so bitmaps of different lengths sit packed and the arena is uploaded as it is.
    void align(const sa_mode &mode, std::vector<sa_problem> &pr, const std::vector<uint32_t> &arena, AlignOut &out) {
        out.pk = sa_packed_paths{};
        if (pr.empty()) return;
        const int rc = sa_align_batch_packed(ctx, &mode, pr.data(), (int64_t)pr.size(), arena.empty() ? nullptr : arena.data(), 1,
                                             (int64_t)arena.size(), &out.pk);
        if (rc) die("sa_align_batch_packed", rc);
    }
    // Device-resident alignment loop (sa_align_rounds): every round of every pair in one call.  out.nr(k) = number of
    // rounds of pair k (-1: the pair did not fit the device scratch, the caller runs it through align()).
    struct RoundsOut {
        sa_rounds_out ro{};
        int nr(size_t k) const { return ro.n_rounds[k]; }
        struct View {
            const sa_round *rec;
            const int32_t *pj, *pq;
            const float *ps;
        };
        View get(size_t k, int r) const {
            const sa_round *rec = ro.rounds + (size_t)ro.round_off[k] + (size_t)r;  // records are packed in pair order
            return View{rec, ro.path_j + rec->path_off, ro.path_q + rec->path_off, ro.path_s + rec->path_off};
        }
    };
    // flt: tracebacks that cannot become a file any more stay on the device (their round record has path_off = -1)
    void align_rounds(const sa_mode &mode, const std::vector<sa_problem> &pr, const std::vector<float> &thr, int max_alns,
                      const std::vector<uint32_t> &arena, const sa_rounds_filter &flt, RoundsOut &out) {
        out.ro = sa_rounds_out{};
        if (pr.empty()) return;
        const int rc = sa_align_rounds_filtered(ctx, &mode, pr.data(), (int64_t)pr.size(), thr.data(), max_alns,
                                                arena.empty() ? nullptr : arena.data(), 1, (int64_t)arena.size(), &flt, &out.ro);
        if (rc) die("sa_align_rounds_filtered", rc);
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
                                       const std::vector<Bitmap> *used_seed, bool tip_aln, float stop_score) {
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
    auto partial_thresh = [&](const PairState &p, int first_lab, int last_lab) -> float {
        const float *contig = contigs.map(p.c);
        const int contig_size = contigs.size(p.c);
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
        // what emit() below tests first, from the path's ends, length and total score alone (the mean again, exactly, there)
        sa_rounds_filter flt;
        flt.tip = tip_aln ? 1 : 0;
        flt.min_len = tip_aln ? min_tip_aln_labels : std::max(min_aln_labels, min_tip_aln_labels);
        flt.mean_min = mean_thresh;
        flt.mean_min_short = 9400.0f;
        // A round that scores below what a file needs ends the pair: later rounds cannot score higher, a round that cannot be
        // written leaves nothing behind but its count, and the count only names files.  (stop_score: see can_write_a_file.)
        flt.stop_below = getenv("SA_NO_EARLY_STOP") ? 0.0f : stop_score;
        flt.tip_windows = (tip_aln && !getenv("SA_NO_TIP_WINDOWS")) ? 1 : 0;  // pairs no end-touching alignment of which can score that much: -2
        eng.align_rounds(mode, rp, rthr, max_alns, arena, flt, ro);
        const double t_r1 = wall_now();
        const size_t n_pairs = st.size();
        // replay + file writing: a worker per 4096 pairs (a standard pass writes about one file per pair)
        const int n_workers = (int)std::max<size_t>(1, std::min<size_t>(std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u),
                                                                        n_pairs / 4096 + 1));
        std::vector<std::vector<std::pair<int, int>>> disc((size_t)n_workers);
        std::vector<int> bad((size_t)n_workers, 0);
        size_t n_fallback = 0;
        auto work = [&](int wi) {
            std::vector<AlnEntry> aln;
            const size_t lo = n_pairs * (size_t)wi / (size_t)n_workers, hi = n_pairs * ((size_t)wi + 1) / (size_t)n_workers;
            for (size_t k = lo; k < hi; k++) {
                PairState &p = st[k];
                const int nr = ro.nr(k);
                if (nr == -2) {  // certified: no round of this pair can write a tip alignment (sa_rounds_filter.tip_windows)
                    p.live = false;
                    continue;
                }
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
                    p.exp_thresh = partial_thresh(p, v.rec->j_first, v.rec->j);
                    p.curr_best = v.rec->score;
                    if (!(p.curr_best > p.exp_thresh)) {  // not accepted: the loop ends (equality would spin in the reference)
                        more = false;
                        r++;
                        break;
                    }
                    p.count += 1;
                    if (v.rec->path_off >= 0) emit(p, len, v.pj, v.pq, v.ps, aln, disc[(size_t)wi]);  // else: certified unwritable
                    more = p.count < max_alns && !(p.curr_best < flt.stop_below);  // the device stopped here too (stop_below)
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
        if (getenv("SA_VERBOSE")) {
            long long n_rounds = 0, n_first_below = 0, n_hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (size_t k = 0; k < n_pairs; k++) {
                const int nr = ro.nr(k);
                if (nr <= 0) continue;
                n_rounds += nr;
                n_hist[std::min(nr, 7)]++;
                n_first_below += ro.get(k, 0).rec->score < flt.stop_below;
            }
            fprintf(stderr, "  rounds: %lld in all (pairs by number of rounds 1..7+: %lld %lld %lld %lld %lld %lld %lld); first round already below %.0f: %lld pairs\n",
                    n_rounds, n_hist[1], n_hist[2], n_hist[3], n_hist[4], n_hist[5], n_hist[6], n_hist[7], (double)flt.stop_below, n_first_below);
        }
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
                p.exp_thresh = partial_thresh(p, pj[len - 1], pj[0]);
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

// One SegAligner run (the body of the reference's main, main.cpp:453-656).  `eng` may already hold an open context
// (resident server); `keep_open` leaves it open on return.  `sp`: this process's place among the GPU processes of the run
// (rank 0 alone prints and writes the score files; every rank writes the alignment files of its own pairs).
static int run_job(int argc, char *argv[], Engine &eng, bool keep_open, Spmd &sp) {
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
    pt.on = pt.on && sp.rank == 0;
    // CUDA start-up (driver init + context, ~0.3-0.6 s for the one device this process sees) overlaps with CMAP parsing.
    // Rank 0 starts it right away; the other ranks wait until they know the run is big enough to need them.
    JobExit open_err{0, false};
    std::thread gpu_open;
    auto start_open = [&]() {
        gpu_open = std::thread([&eng, &open_err, &sp]() {
            try {
                if (!eng.ctx) sp.wait_init_turn();
                eng.open();
            } catch (const JobExit &e) {
                open_err = e;
            }
            sp.pass_init_turn();  // also when it failed: the ranks behind must not wait for ever
        });
    };
    struct Joiner {  // the opener thread must be joined on every exit path (exceptions included)
        std::thread &t;
        ~Joiner() {
            if (t.joinable()) t.join();
        }
    } joiner{gpu_open};
    if (sp.rank == 0) start_open();
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

    // GPUs: every process that was forked for big runs, one otherwise (a small run is over before a second context is up).
    // Every rank takes the same decision from the same inputs; the ranks that are not needed leave before touching CUDA.
    double total_cells = 0;
    {
        double sl = 0, cl = 0;
        for (int m = 0; m < segs.n_maps(); m++) sl += segs.labels(m);
        for (int m = 0; m < contigs.n_maps(); m++) cl += contigs.labels(m);
        total_cells = sl * cl;
    }
    {
        int asked = o.n_gpus;
        if (asked <= 0)
            if (const char *e = getenv("SA_NGPUS")) asked = atoi(e);
        if (asked <= 0) {
            // Nobody asked for a number of GPUs: a fixed-size job gains from another GPU only while its share of the work
            // outweighs the start-up ladder (0.8 - 2 s per additional GPU, driver-side and box-dependent; every rank must be
            // up before the first exchange, and a rank that is already scoring stalls in its driver calls while the others
            // come up).  Whole run ~ cells / 38 GCUPS on one GPU; f(N) = 1.2 (N - 1) + t1 / N; the smallest N within 5 % of the
            // minimum (the curve is flat there, and fewer processes is the safer side of a start-up cost that varies).
            const double t1 = total_cells / 38e9;
            auto f = [&](int n) { return 1.2 * (n - 1) + t1 / n; };
            int best = 1;
            for (int n = 2; n <= sp.world; n++)
                if (f(n) < f(best)) best = n;
            for (int n = 1; n < best; n++)
                if (f(n) <= 1.05 * f(best)) {
                    best = n;
                    break;
                }
            sp.world = std::max(1, std::min(sp.world, best));
        }
        if (o.fitting_aln || segs.n_maps() == 0 || contigs.n_maps() == 0) sp.world = 1;
    }
    if (sp.rank >= sp.world) {
        sp.pass_init_turn();
        return 0;
    }
    if (sp.rank > 0) start_open();

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
    gpu_open.join();
    if (open_err.code) throw open_err;
    pt.lap("open GPU context");
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
        // this rank scores its shard of the problem list; exchange 1: every rank needs every score for the
        // thresholds, so the shards meet in the shared region (a private vector when the run has one process)
        const std::vector<int64_t> mine = my_shard(pr, contigs, segs, sp.world, sp.rank);
        std::vector<sa_problem> mp(mine.size());
        for (size_t k = 0; k < mine.size(); k++) mp[k] = pr[(size_t)mine[k]];
        if (o.detection) {
            int total_labels = 0;
            for (int c = 0; c < contigs.n_maps(); c++) total_labels += contigs.size(c);  // main.cpp:118-121
            std::vector<int64_t> off(pr.size() + 1, 0);  // global slots: n_want scores per problem
            for (size_t k = 0; k < pr.size(); k++)
                off[k + 1] = off[k] + int(roundf(o.n_detect_scores * float(contigs.size(pr[k].contig)) / float(total_labels))) + 1;
            std::vector<float> own_sc;
            std::vector<int32_t> own_got;
            float *all_sc;
            int32_t *all_got;
            if (sp.world > 1) {
                all_sc = sp.alloc<float>((size_t)off[pr.size()]);
                all_got = sp.alloc<int32_t>(pr.size());
            } else {
                own_sc.resize((size_t)off[pr.size()]);
                own_got.resize(pr.size());
                all_sc = own_sc.data();
                all_got = own_got.data();
            }
            std::vector<int32_t> want(mp.size()), got;
            for (size_t k = 0; k < mp.size(); k++) want[k] = (int32_t)(off[(size_t)mine[k] + 1] - off[(size_t)mine[k]]);
            std::vector<int64_t> loff;
            std::vector<float> sc;
            eng.detect(mode, mp, want, loff, sc, got);
            for (size_t k = 0; k < mp.size(); k++) {
                std::copy(sc.begin() + loff[k], sc.begin() + loff[k] + got[k], all_sc + off[(size_t)mine[k]]);
                all_got[(size_t)mine[k]] = got[k];
            }
            sp.barrier();
            for (size_t k = 0; k < pr.size(); k++)
                for (int32_t t = 0; t < all_got[k]; t++)
                    rows.push_back(ScoreRow{segs.ids[(size_t)pr[k].seg], contigs.ids[(size_t)pr[k].contig], all_sc[off[k] + t]});
        } else {
            // While the GPU scores, a helper thread pins the host buffers the alignment passes will read their results from:
            // pinning costs 0.3 - 1 s per GB, and growing a pinned buffer later means freeing the old one, which waits for the
            // device.  Sized for 60 % of this rank's pairs reaching the tip pass (C2: 58 %, C5: 38 %); more just grows later.
            std::thread prewarm;
            struct Joiner {
                std::thread &t;
                ~Joiner() {
                    if (t.joinable()) t.join();
                }
            } prewarm_joiner{prewarm};
            if (!o.local_aln && !o.swap_b_x && o.lookback == 6 && !getenv("SA_NO_ROUNDS"))
                prewarm = std::thread([&eng, &o, n_hint = (int64_t)(0.6 * (double)mine.size()) + 1024]() {
                    (void)sa_prepare_rounds(eng.ctx, n_hint, std::max(o.max_seg_contig_alns, 6));
                });
            std::vector<float> own;
            float *all;
            if (sp.world > 1) all = sp.alloc<float>(pr.size());
            else {
                own.resize(pr.size());
                all = own.data();
            }
            std::vector<sa_result> res;
            if (sp.world <= 1) {
                eng.score(mode, mp, res);
                for (size_t k = 0; k < mp.size(); k++) all[(size_t)mine[k]] = res[k].score;
            } else {
                // Dynamic units instead of the static shard: the ranks come up one after the other (Spmd::wait_init_turn), so
                // each takes the next run of the shared largest-first order whenever it is free -- half an even share of what is
                // left, at least ~40 ms of scoring -- through one atomic cursor in the exchange region.  Scores land by problem
                // index, so the result does not depend on who computed what.
                std::vector<int64_t> order;
                std::vector<double> cum;
                size_class_order(pr, contigs, segs, order, cum);
                const long long n_all = (long long)pr.size();
                const double total = cum.back(), min_unit = std::max(total / (64.0 * sp.world), 2.0e9);
                int n_units = 0;
                double my_cells = 0;
                for (;;) {
                    long long b = sp.ctl->next_unit.load(), e = b;
                    for (;;) {
                        if (b >= n_all) break;
                        const double want = std::max(min_unit, (total - cum[(size_t)b]) / (2.0 * sp.world));
                        e = (long long)(std::upper_bound(cum.begin() + b + 1, cum.end(), cum[(size_t)b] + want) - cum.begin());
                        e = std::max(std::min(e, n_all), b + 1);
                        if (sp.ctl->next_unit.compare_exchange_weak(b, e)) break;
                    }
                    if (b >= n_all) break;
                    mp.resize((size_t)(e - b));
                    for (long long k = b; k < e; k++) mp[(size_t)(k - b)] = pr[(size_t)order[(size_t)k]];
                    eng.score(mode, mp, res);
                    for (long long k = b; k < e; k++) all[(size_t)order[(size_t)k]] = res[(size_t)(k - b)].score;
                    n_units++;
                    my_cells += cum[(size_t)e] - cum[(size_t)b];
                }
                if (getenv("SA_VERBOSE"))
                    fprintf(stderr, "  rank %d scored %d units, %.3g of %.3g cells, done %.3f s after start\n", sp.rank, n_units, my_cells, total,
                            wall_now() - w0);
            }
            sp.barrier();
            rows.reserve(pr.size());
            for (size_t k = 0; k < pr.size(); k++)
                rows.push_back(ScoreRow{segs.ids[(size_t)pr[k].seg], contigs.ids[(size_t)pr[k].contig], all[k]});
        }
    }
    pt.lap("phase 1 scoring (GPU)");
    // *_all_scores.txt (10^6..10^7 rows) is formatted and written beside the thresholds and the alignment passes, which only
    // read `rows`; joined before the run reports its end
    std::thread scores_writer;
    struct ScoresJoiner {
        std::thread &t;
        ~ScoresJoiner() {
            if (t.joinable()) t.join();
        }
    } scores_joiner{scores_writer};
    if (sp.rank == 0) scores_writer = std::thread([&rows, &o]() { write_all_scores(rows, o.sample_prefix); });

    // ---------------- phase 2: thresholds (main.cpp:555-557); every rank computes the same values ---------------------
    std::vector<std::vector<std::pair<int, float>>> thr2;
    compute_score_thresholds_multi(rows, segs, contigs, {(double)o.p_val, (double)o.p_val_tip}, o.n_detect_scores, thr2);
    const auto &score_thresholds = thr2[0];
    const auto &tip_thresholds = thr2[1];
    printf("Writing scoring distribution.\n");
    if (sp.rank == 0) write_score_thresholds(score_thresholds, o.sample_prefix + o.detection_label);
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
    // Pairs that cannot produce an alignment file are not aligned at all.  A file needs >= 5 labels (tip, -gen=1: >= 3) and a
    // mean score per step >= 8750 (7000 in the standard pass; 9400 below 5 labels), i.e. a round's best score >= 35000 (18800).
    // Every round's best score is <= the pair's phase-1 score: both are the maximum of the same semi-global DP, and masking
    // contig labels only removes candidates (each cell is a max over predecessor + constant; fp32 '+' and max are monotone).
    // So a pair whose phase-1 score is below that bound can neither write a file nor contribute used labels, in any round;
    // whether the reference grinds through its rounds or not is unobservable.  Not for -local / -detection, whose alignment
    // rounds use a different initialisation than phase 1.  (1 % slack for the rounding of the mean's division.)
    std::vector<float> phase1;  // by (segment idx, contig idx); +inf = unknown
    float min_file_score = 0.0f;
    if (!o.local_aln && !getenv("SA_NO_SKIP")) {
        phase1.assign((size_t)segs.n_maps() * (size_t)contigs.n_maps(), INFINITY);
        for (size_t k = 0; k < pr.size(); k++) phase1[(size_t)pr[k].seg * (size_t)contigs.n_maps() + (size_t)pr[k].contig] = rows[k].score;
    }
    auto can_write_a_file = [&](const std::vector<std::pair<int, int>> &all) {
        if (phase1.empty()) return all;
        std::vector<std::pair<int, int>> out;
        out.reserve(all.size());
        for (auto &e : all)
            if (!(phase1[(size_t)e.first * (size_t)contigs.n_maps() + (size_t)e.second] < min_file_score)) out.push_back(e);
        if (getenv("SA_VERBOSE") && sp.rank == 0)
            fprintf(stderr, "  %zu of %zu pairs can reach the score an alignment file needs (%.0f); the others are skipped\n", out.size(),
                    all.size(), (double)min_file_score);
        return out;
    };
    // this rank's share of a pair list (same deterministic partition on every rank)
    auto my_pairs = [&](const std::vector<std::pair<int, int>> &all) {
        if (sp.world <= 1) return all;
        std::vector<sa_problem> as_pr(all.size());
        for (size_t k = 0; k < all.size(); k++) as_pr[k] = sa_problem{all[k].second, all[k].first, -1, 0};
        std::vector<std::pair<int, int>> out;
        for (int64_t k : my_shard(as_pr, contigs, segs, sp.world, sp.rank)) out.push_back(all[(size_t)k]);
        return out;
    };
    printf("Performing alignments\n");
    fflush(stdout);
    min_file_score = 0.99f * 35000.0f;  // standard pass: >= 6 labels, mean >= 7000 (main.cpp:218-222,249-255,310-325)
    std::vector<Bitmap> contig_used =
        run_alignment_pass(eng, o, segs, contigs, my_pairs(can_write_a_file(uniq)), score_thresholds, nullptr, false, phase1.empty() ? 0.0f : min_file_score);
    if (sp.world > 1) {  // exchange 2 (main.cpp:597-602): union of the labels of the written alignments
        size_t words = 0;
        for (auto &b : contig_used) words += b.size();
        uint32_t *all = sp.alloc<uint32_t>(words * (size_t)sp.world);
        uint32_t *own = all + words * (size_t)sp.rank;
        for (auto &b : contig_used) {
            std::copy(b.begin(), b.end(), own);
            own += b.size();
        }
        sp.barrier();
        for (int r = 0; r < sp.world; r++) {
            const uint32_t *src = all + words * (size_t)r;
            for (auto &b : contig_used)
                for (auto &w : b) w |= *src++;
        }
    }
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
        if (getenv("SA_VERBOSE") && sp.rank == 0) {  // how much of the tip pass's first sweep the used labels could save
            long long rows = 0, used = 0, strips = 0, full = 0, blocks = 0;
            for (int c : tip_contigs) {
                const Bitmap &b = contig_used[(size_t)c];
                const int nb = contigs.labels(c);
                rows += nb;
                strips += (nb + 31) / 32;
                for (int s = 0; s * 32 < nb; s++) {
                    const int r = std::min(32, nb - s * 32);
                    const uint32_t valid = r >= 32 ? 0xffffffffu : ((1u << r) - 1u);
                    full += (b[(size_t)s] & valid) == valid;
                }
                for (int j = 0; j < nb;) {
                    if ((b[(size_t)(j >> 5)] >> (j & 31)) & 1u) {
                        used++;
                        j++;
                        continue;
                    }
                    blocks++;
                    for (int k = j; k < std::min(nb, j + 32); k++) used += (b[(size_t)(k >> 5)] >> (k & 31)) & 1u;
                    j += 32;
                }
            }
            fprintf(stderr, "  tip pass: %zu contigs, %lld labels, %lld used; %lld aligned strips, %lld of them entirely used; %lld blocks of 32 rows if a block starts at the next free label\n",
                    tip_contigs.size(), rows, used, strips, full, blocks);
        }
        printf("Performing alignments\n");
        fflush(stdout);
        min_file_score = 0.99f * ((o.inst_gen == 1) ? 2.0f * 9400.0f : 4.0f * 8750.0f);  // tip: >= 3 labels at 9400 / >= 5 at 8750
        run_alignment_pass(eng, o, segs, contigs, my_pairs(can_write_a_file(tip_pairs)), tip_thresholds, &contig_used, true, phase1.empty() ? 0.0f : min_file_score);
        printf("Finished tip alignments.\n\n");
        pt.lap("phase 4 tip alignment");
    }
    {   // exchange 3: the run is over when every rank has written its files
        uint64_t h2d = 0, d2h = 0;
        sa_get_traffic(eng.ctx, &h2d, &d2h);
        if (sp.world > 1) {
            sp.ctl->h2d.fetch_add(h2d);
            sp.ctl->d2h.fetch_add(d2h);
            sp.barrier();
            h2d = sp.ctl->h2d.load();
            d2h = sp.ctl->d2h.load();
        }
        if (scores_writer.joinable()) scores_writer.join();
        printf("elapsed seconds for run %f\n", double(clock() - begin) / CLOCKS_PER_SEC);
        if (getenv("SA_VERBOSE") && sp.rank == 0) {
            fprintf(stderr, "SegAligner(b200): copied H2D %llu bytes, D2H %llu bytes\n", (unsigned long long)h2d, (unsigned long long)d2h);
            fprintf(stderr, "SegAligner(b200): wall %.3f s on %d GPU(s)\n", wall_now() - w0, sp.world);
        }
    }
    fflush(stdout);
    fflush(stderr);
    if (sp.world > 1 && sp.rank > 0) {
        // A worker rank is done once every rank has passed the last exchange: its files are written.  Tearing its context down
        // (serialised by the driver across the processes of a box, ~0.4 s each) happens at process exit, in the background:
        // give up the inherited stderr so that a caller reading our pipes does not wait for it.
        const int dn = ::open("/dev/null", O_WRONLY);
        if (dn >= 0) {
            dup2(dn, 2);
            ::close(dn);
        }
        return 0;
    }
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
#include <sys/prctl.h>
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

// The GPUs this run may use, as CUDA_VISIBLE_DEVICES tokens, WITHOUT initialising CUDA (that is the expensive step a
// multi-GPU run wants to do once per process, in parallel): the caller's own CUDA_VISIBLE_DEVICES list, else one entry
// per /dev/nvidia<N> device node.  Empty when neither exists (CUDA then reports the absence of a device itself).
std::vector<std::string> visible_devices() {
    std::vector<std::string> out;
    if (const char *vis = getenv("CUDA_VISIBLE_DEVICES")) {
        std::string t;
        for (const char *c = vis;; c++) {
            if (*c == ',' || *c == 0) {
                if (!t.empty()) out.push_back(t);
                t.clear();
                if (*c == 0) break;
            } else {
                t += *c;
            }
        }
        return out;
    }
    for (int n = 0; n < 64; n++) {
        struct stat sb;
        if (stat(("/dev/nvidia" + std::to_string(n)).c_str(), &sb) == 0) out.push_back(std::to_string((int)out.size()));
    }
    return out;
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
    // The server drives ONE GPU: hiding the other devices before the first CUDA call saves seconds of driver
    // initialisation on a multi-GPU box (big multi-GPU runs are not what a resident server is for).
    {
        const std::vector<std::string> devs = visible_devices();
        if (!devs.empty()) setenv("CUDA_VISIBLE_DEVICES", devs[0].c_str(), 1);
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
            Spmd solo;  // the server drives one GPU from one process
            code = (uint32_t)run_job((int)args.size(), av.data(), eng, true, solo);
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
    int asked = 0;  // -ngpus=N, else SA_NGPUS, else 0 = all visible GPUs (big runs) / one (small runs), decided after parsing
    for (int i = 3; i < argc; i++) {
        fitting = fitting || strcmp(argv[i], "-fitting") == 0;
        if (strncmp(argv[i], "-ngpus=", 7) == 0) asked = atoi(argv[i] + 7);
    }
    if (asked <= 0)
        if (const char *e = getenv("SA_NGPUS")) asked = atoi(e);
    if (forced || (fitting && !forbidden) || (argc == 2 && strcmp(argv[1], "--server-stop") == 0)) {
        const int rc = run_via_server(argc, argv, forced ? 300 : 30);
        if (rc >= 0) return rc;
    }
    // One process per GPU, forked before the first CUDA call (see Spmd).  SA_GPU_OVERSUBSCRIBE=1 (testing aid): -ngpus=N may
    // exceed the visible devices; rank r then runs on device r % visible, which exercises the same sharded code on one GPU.
    const std::vector<std::string> devs = visible_devices();
    const int n_dev = (int)devs.size();
    const char *ov = getenv("SA_GPU_OVERSUBSCRIBE");
    const bool oversubscribe = ov && atoi(ov) != 0;
    int n_proc = 1;
    if (n_dev > 0 && argc >= 3 && !fitting)
        n_proc = asked > 0 ? (oversubscribe ? std::min(asked, 64) : std::min(asked, n_dev)) : n_dev;
    Spmd sp;
    sp.world = n_proc;
    if (n_proc > 1) {
        // exchange region: virtual size only (pages appear when touched); scores of 10^9 problems would fit
        for (size_t bytes = (size_t)16 << 30; bytes >= ((size_t)64 << 20) && !sp.region; bytes >>= 2) {
            void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (m != MAP_FAILED) {
                sp.region = (char *)m;
                sp.region_bytes = bytes;
            }
        }
        if (!sp.region) {
            sp.world = n_proc = 1;
        } else {
            sp.ctl = new (sp.region) SpmdCtl();
            sp.ctl->arrived.store(0);
            sp.ctl->generation.store(0);
            sp.ctl->error.store(0);
            sp.ctl->h2d.store(0);
            sp.ctl->d2h.store(0);
            sp.ctl->init_turn.store(0);
            sp.ctl->next_unit.store(0);
            fflush(stdout);
            fflush(stderr);
            for (int r = 1; r < n_proc; r++) {
                const pid_t pid = fork();
                if (pid == 0) {
                    prctl(PR_SET_PDEATHSIG, SIGTERM);
                    sp.rank = r;
                    sp.kids.clear();
                    const int dn = open("/dev/null", O_WRONLY);  // rank 0 alone reports progress on stdout
                    if (dn >= 0) {
                        dup2(dn, 1);
                        close(dn);
                    }
                    break;
                }
                if (pid < 0) {  // could not fork: the ranks that exist would wait for this one forever
                    sp.ctl->error.store(1);
                    fprintf(stderr, "SegAligner(b200): fork failed\n");
                    return 2;
                }
                sp.kids.push_back(pid);
            }
        }
    }
    if (n_dev > 0) setenv("CUDA_VISIBLE_DEVICES", devs[(size_t)(sp.rank % n_dev)].c_str(), 1);
    Engine eng;
    int code = 0;
    try {
        code = run_job(argc, argv, eng, false, sp);
    } catch (const JobExit &e) {
        code = e.code;
        if (sp.ctl) sp.ctl->error.store(1);  // releases the other ranks from their barrier
    }
    fflush(stdout);
    fflush(stderr);
    if (sp.rank > 0) _exit(code);
    // The run is over when every rank's files are written.  After a successful run every rank has passed the last exchange, so
    // the workers only have their CUDA teardown left: they are not waited for (they are reaped by init).  After a failure the
    // workers are collected so that nothing of this run is left behind.
    for (pid_t k : sp.kids) {
        int st = 0;
        if (code == 0 && sp.world > 1) {
            if (waitpid(k, &st, WNOHANG) == k && !(WIFEXITED(st) && WEXITSTATUS(st) == 0)) code = 2;
        } else if (waitpid(k, &st, 0) == k && !(WIFEXITED(st) && WEXITSTATUS(st) == 0) && code == 0 && sp.world > 1) {
            code = 2;
        }
    }
    return code;
}
