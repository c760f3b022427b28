// C-ABI host layer of libsegaligner_b200.so (declared in include/segaligner_b200.h).
// Owns device memory, the stream, batching/chunking and the largest-first problem order; all arithmetic that
// decides results runs in the kernels of sa_kernels.cu, except the per-label operand tables built here on the
// host in exact fp32 (same operations, same order as score_f, /root/reference/SegAligner/SegAligner.cpp:86-90).
// There is NO CPU fallback: every entry point fails with SA_ENODEV / SA_ECUDA when no sm_100 GPU is usable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>
#include "sa_kernels.cuh"

using namespace sa;

int sa_digest_run(cudaStream_t st, int n_sms, const uint8_t *d_seq, int64_t n, const char *motif, int32_t motif_len, int32_t offset,
                  long long *d_out, unsigned long long *d_count, int64_t cap);

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();  // clear the sticky allocation error, or the next launch check would report it
            e = cudaMalloc(&p, bytes);
            want = bytes;
            if (e != cudaSuccess) (void)cudaGetLastError();
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

// grow-only pinned host buffer (results / packed paths travel D2H at full PCIe speed, no page faults per call)
struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes, bool keep = false) {
        if (bytes <= cap) return cudaSuccess;
        void *np = nullptr;
        const size_t want = keep ? std::max(bytes + bytes / 4 + 4096, 2 * cap) : bytes + bytes / 4 + 4096;  // appended-to buffers double
        cudaError_t e = cudaMallocHost(&np, want);
        if (e != cudaSuccess) return e;
        if (keep && p && cap) memcpy(np, p, cap);
        if (p) cudaFreeHost(p);
        p = np;
        cap = want;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

struct MapSetHost {
    std::vector<int64_t> off;  // n_maps+1
    DevBuf d_pos, d_prob, d_ops, d_off;
    int n_maps = 0;
    int max_labels = 0;
    MapSetDev dev() const {
        MapSetDev m;
        m.pos = d_pos.as<float>();
        m.prob = d_prob.as<float>();
        m.ops = d_ops.as<LabelOps>();
        m.pos_off = d_off.as<int64_t>();
        m.n_maps = n_maps;
        m.max_labels = max_labels;
        return m;
    }
    int labels(int m) const { return (int)(off[m + 1] - off[m]) - 1; }
};

}  // namespace

struct sa_ctx {
    int device = 0;
    int n_sms = 0;
    int ctas_per_sm = 1;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // cooperative (large-problem) launches run beside the bulk kernel
    cudaStream_t stream3 = nullptr;  // cluster-cooperative (huge-problem) launches
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join3 = nullptr;
    int n_clusters = 0;              // co-resident clusters of the cluster variant (0 = unavailable)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_user[2] = {nullptr, nullptr};
    std::string err;
    MapSetHost sides[2];
    DevBuf d_problems, d_order, d_results, d_halo, d_counter, d_cell_off, d_S, d_code, d_used, d_path_off, d_path_j,
        d_path_q, d_path_s, d_tmp0, d_tmp1, d_tmp2, d_halo2, d_row_off, d_rowmax_s, d_rowmax_q, d_prev32, d_halo3, d_mail;
    HostBuf h_res, h_start, h_pj, h_pq, h_ps;  // packed alignment output of the last sa_align_batch* call
    // sa_align_rounds: per-worker scratch that persists across the rounds of a pair, the packed path arena, its host mirrors
    DevBuf d_r_code, d_r_halo, d_r_col2, d_r_lastrow, d_r_used, d_r_path, d_r_ps, d_r_thr, d_r_rounds, d_r_nrounds, d_r_roff, d_r_packed, d_r_pj, d_r_pq,
        d_r_psc, d_r_ctl;
    DevBuf d_units, d_mbox, d_cpl_res, d_cpl_done;  // coupled units: queue, mailboxes, per-unit start candidates, per-problem counters
    HostBuf h_cpl_err;                              // the handshake-timeout flag of the last coupled launch lands here
    bool cpl_pending = false;
    uint32_t cpl_epoch = 0;
    HostBuf h_stage;                    // pinned D2H landing zone of a chunk's round records
    std::vector<int32_t> h_nr;          // what sa_rounds_out points at (scatter targets: pageable memory is enough)
    std::vector<sa_round> h_rounds;     // packed, pair order
    std::vector<int64_t> h_roff;
    sa_timing timing = {0, 0, 0, 0};
    size_t store_budget_cells = 0;  // max cells of stored matrices per chunk
    uint64_t h2d_bytes = 0, d2h_bytes = 0;  // bytes moved by this context since it was created (sa_get_traffic)
};

// every host<->device copy of the library goes through here so that the traffic of a run can be reported exactly
static inline cudaError_t cpy(sa_ctx *ctx, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
    if (kind == cudaMemcpyHostToDevice) ctx->h2d_bytes += bytes;
    else if (kind == cudaMemcpyDeviceToHost) ctx->d2h_bytes += bytes;
    return cudaMemcpyAsync(dst, src, bytes, kind, st);
}

struct sa_plan;

namespace {

int fail(sa_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code;
}
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? SA_ENOMEM : SA_ECUDA,                   \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                           \
    } while (0)

// score_f's operands per label, exact fp32, reference operation order.
void build_label_ops(const float *pos, const float *prob, int n, LabelOps *out) {
    for (int l = 0; l < n; l++) {
        for (int k = 1; k <= LB; k++) {
            float d = 0.0f, f = 0.0f;
            if (l - k >= 0) {
                volatile float dd = pos[l] - pos[l - k];  // b_posns[j] - b_posns[i]  (SegAligner.cpp:87)
                d = dd;
                volatile float acc = 0.0f;  // std::accumulate(probs+p+1, probs+q, 0.0f)  (:86)
                for (int t = l - k + 1; t < l; t++) acc = acc + prob[t];
                volatile float ff = 5000.0f * acc;  // fp_t (:90)
                f = ff;
            }
            out[l].d[k - 1] = d;
            out[l].f[k - 1] = f;
        }
    }
}

// Processing order.  Problems are split into two classes:
//   large  -> cooperative kernel (one CTA of COOP_WARPS warps per problem), largest first;
//   bulk   -> one warp per problem, largest first (LPT) via a counting sort on a 1/8-octave size class, O(n).
// A problem is "large" when one warp would need longer for it than the whole batch needs on the whole GPU
// (cells_p * n_warp_slots > alpha * total_cells), so that no single problem can become the tail.
//   huge   -> cluster-cooperative kernel (CLUSTER_CTAS SMs per problem) when even ONE SM would need longer for
//             the problem than the whole batch needs on the whole GPU (cells_p * n_sms > alpha * total_cells).
// Since round 2 the default for problems that are too large for one warp is "coupled units" (SA_SHAPES=coupled, the
// default): the problem is cut into U units of row strips that any U warps sweep as a wavefront, handing rows over through
// tagged mailboxes in L2 (see fill_kernel, CPL).  A problem is coupled when one warp would need more than cpl_alpha x the
// batch's ideal makespan for it; U is chosen so that a unit is half of that (and U <= 64, <= its strips, <= nx / 2 so the ring
// of units never waits for its own tail).  SA_SHAPES=legacy keeps the two barrier-synchronised shapes above.
struct Plan {
    std::vector<int32_t> order;  // bulk problems first, then the large ones, then the huge ones (or the coupled ones)
    int64_t n_bulk = 0, n_large = 0, n_huge = 0, n_cpl = 0;
    std::vector<CplUnit> units;  // coupled mode: the unit queue, largest problems first, units of a problem consecutive
    int cpl_max_nx = 0;
    double cpl_max_unit_cells = 0;  // largest unit (a unit's warp may legitimately be waited for that long)
    double total_cells = 0;
    // every problem, largest classes first (one-warp-per-problem consumers: the rounds kernel)
    std::vector<int32_t> big_first() const {
        std::vector<int32_t> o;
        o.reserve(order.size());
        o.insert(o.end(), order.begin() + n_bulk + n_large, order.end());
        o.insert(o.end(), order.begin() + n_bulk, order.begin() + n_bulk + n_large);
        o.insert(o.end(), order.begin(), order.begin() + n_bulk);
        return o;
    }
};

void build_plan(const sa_ctx *c, const sa_problem *pr, int64_t n, int n_warp_slots, Plan &plan) {
    const int NB = 8 * 40;
    std::vector<int64_t> cnt(NB + 1, 0);
    std::vector<uint16_t> cls((size_t)n);
    std::vector<double> cells((size_t)n);
    double total = 0;
    for (int64_t k = 0; k < n; k++) {
        cells[(size_t)k] = (double)c->sides[0].labels(pr[k].contig) * (double)c->sides[1].labels(pr[k].seg);
        total += cells[(size_t)k];
    }
    bool legacy = false;
    if (const char *e = getenv("SA_SHAPES")) legacy = strcmp(e, "legacy") == 0;
    plan.units.clear();
    plan.n_cpl = 0;
    plan.cpl_max_nx = 0;
    plan.cpl_max_unit_cells = 0;
    if (!legacy) {
        double a = 0.25;
        if (const char *e = getenv("SA_CPL_ALPHA")) a = atof(e);
        const double makespan = total / (double)n_warp_slots;
        struct Big {
            double neg_cells;
            int32_t k, U;
        };
        std::vector<Big> big;
        double unit_frac = 0.5, u_max = 64.0;  // a unit is this fraction of the coupling threshold; at most this many units
        if (const char *e = getenv("SA_CPL_UNIT")) unit_frac = atof(e);
        if (const char *e = getenv("SA_CPL_UMAX")) u_max = std::max(2.0, std::min(64.0, atof(e)));
        // (a unit is at most ~1 s of one warp: waits on units that are not running yet stay short; the loop bounds the mailbox memory)
        for (double target = std::min(std::max(a * makespan * unit_frac, 2048.0), 1.3e7);; target *= 2.0) {
            big.clear();
            std::fill(cnt.begin(), cnt.end(), 0);
            int64_t n_units = 0;
            for (int64_t k = 0; k < n; k++) {
                const double cl = cells[(size_t)k];
                const int nb = c->sides[0].labels(pr[k].contig), nx = c->sides[1].labels(pr[k].seg);
                const int nstrips = (nb + 31) / 32;
                int U = 0;
                if (a > 0 && cl > a * makespan && nstrips >= 2 && nx >= 4 && (double)nstrips * nx < 268435456.0)  // 28-bit step tags
                    U = (int)std::min<double>(std::min<double>(std::ceil(cl / target), u_max), (double)std::min(nstrips, std::max(2, nx / 4)));  // a ring of U units lags 3 U columns: keep it below nx
                if (U >= 2) {
                    big.push_back(Big{-cl, (int32_t)k, U});
                    cls[(size_t)k] = 0xffff;
                    n_units += U;
                    continue;
                }
                int b = cl > 1.0 ? (int)(std::log2(cl) * 8.0) : 0;
                if (b >= NB) b = NB - 1;
                cls[(size_t)k] = (uint16_t)(NB - 1 - b);  // descending size
                cnt[NB - 1 - b + 1]++;
            }
            if (n_units <= 8ll * n_warp_slots) break;
        }
        for (int b = 0; b < NB; b++) cnt[b + 1] += cnt[b];
        plan.n_large = plan.n_huge = 0;
        plan.n_cpl = (int64_t)big.size();
        plan.n_bulk = n - plan.n_cpl;
        plan.total_cells = total;
        plan.order.resize((size_t)n);
        for (int64_t k = 0; k < n; k++)
            if (cls[(size_t)k] != 0xffff) plan.order[(size_t)cnt[cls[(size_t)k]]++] = (int32_t)k;
        std::sort(big.begin(), big.end(), [](const Big &x, const Big &y) { return x.neg_cells < y.neg_cells || (x.neg_cells == y.neg_cells && x.k < y.k); });
        int32_t slot = 0;
        for (size_t i = 0; i < big.size(); i++) {
            plan.order[(size_t)plan.n_bulk + i] = big[i].k;
            for (int32_t u = 0; u < big[i].U; u++) plan.units.push_back(CplUnit{big[i].k, u, big[i].U, slot});
            slot += big[i].U;
            plan.cpl_max_nx = std::max(plan.cpl_max_nx, c->sides[1].labels(pr[big[i].k].seg));
            plan.cpl_max_unit_cells = std::max(plan.cpl_max_unit_cells, -big[i].neg_cells / big[i].U);
        }
        return;
    }
    double alpha = 0.5;
    if (const char *e = getenv("SA_COOP_ALPHA")) alpha = atof(e);
    const double thresh = alpha * total / (double)n_warp_slots;
    bool no_cluster = false;
    if (const char *e = getenv("SA_NO_CLUSTER")) no_cluster = atoi(e) != 0;
    double huge_min_cells = 4096.0 * 64.0;
    if (const char *e = getenv("SA_HUGE_MIN_CELLS")) huge_min_cells = atof(e);
    // the cluster wavefront pays a cluster barrier per step and idles when nx < its 128 warps, so it is reserved for
    // problems that would otherwise be the tail: single-SM time > alpha_huge x the batch's ideal makespan
    double alpha_huge = 0.7;  // measured on the C3 detection sweep (alignment rounds of ~80 chromosome x region problems)
    if (const char *e = getenv("SA_HUGE_ALPHA")) alpha_huge = atof(e);
    const double thresh_huge =
        (c->n_clusters > 0 && alpha > 0 && !no_cluster) ? std::min(alpha, alpha_huge) * 0 + (alpha < 1e-6 ? alpha : alpha_huge) * total / (double)c->n_sms : 1e300;
    std::vector<std::pair<double, int32_t>> large, huge;
    for (int64_t k = 0; k < n; k++) {
        const double cl = cells[(size_t)k];
        if (cl > thresh_huge && cl >= huge_min_cells) {
            huge.emplace_back(-cl, (int32_t)k);
            cls[(size_t)k] = 0xffff;
            continue;
        }
        if (alpha > 0 && cl > thresh && cl >= 32.0 * 64.0) {
            large.emplace_back(-cl, (int32_t)k);
            cls[(size_t)k] = 0xffff;
            continue;
        }
        int b = cl > 1.0 ? (int)(std::log2(cl) * 8.0) : 0;
        if (b >= NB) b = NB - 1;
        cls[(size_t)k] = (uint16_t)(NB - 1 - b);  // descending size
        cnt[NB - 1 - b + 1]++;
    }
    for (int b = 0; b < NB; b++) cnt[b + 1] += cnt[b];
    plan.n_large = (int64_t)large.size();
    plan.n_huge = (int64_t)huge.size();
    plan.n_bulk = n - plan.n_large - plan.n_huge;
    plan.total_cells = total;
    plan.order.resize((size_t)n);
    for (int64_t k = 0; k < n; k++)
        if (cls[(size_t)k] != 0xffff) plan.order[(size_t)cnt[cls[(size_t)k]]++] = (int32_t)k;
    std::sort(large.begin(), large.end());
    std::sort(huge.begin(), huge.end());
    for (size_t k = 0; k < large.size(); k++) plan.order[(size_t)plan.n_bulk + k] = large[k].second;
    for (size_t k = 0; k < huge.size(); k++) plan.order[(size_t)(plan.n_bulk + plan.n_large) + k] = huge[k].second;
}

}  // namespace

struct sa_plan {
    int64_t n = 0;
    Plan plan;
    DevBuf d_problems, d_order, d_results;
};

namespace {

int check_problems(sa_ctx *ctx, const sa_problem *pr, int64_t n, int64_t n_used_slots) {
    const int nc = ctx->sides[0].n_maps, ns = ctx->sides[1].n_maps;
    if (n >= 2147483648ll) return fail(ctx, SA_EINVAL, "more than 2^31-1 problems in one batch (32-bit order arrays)");
    for (int64_t k = 0; k < n; k++) {
        if (pr[k].contig < 0 || pr[k].contig >= nc || pr[k].seg < 0 || pr[k].seg >= ns)
            return fail(ctx, SA_EINVAL, "problem " + std::to_string(k) + ": map index out of range");
        if (pr[k].used_slot >= n_used_slots) return fail(ctx, SA_EINVAL, "problem: used_slot out of range");
        if (ctx->sides[0].labels(pr[k].contig) < 1 || ctx->sides[1].labels(pr[k].seg) < 1)
            return fail(ctx, SA_EINVAL, "problem " + std::to_string(k) + ": map without labels");
        if ((double)ctx->sides[0].labels(pr[k].contig) * (double)std::max(ctx->sides[1].labels(pr[k].seg), 128) > 34359738368.0)
            return fail(ctx, SA_EINVAL, "problem " + std::to_string(k) + ": more than 2^35 DP cells (32-bit step counter)");
    }
    return SA_OK;
}

int check_unbanded_sizes(sa_ctx *ctx, const sa_mode *mode, const sa_problem *pr, int64_t n) {
    if (!mode || mode->lookback == LB) return SA_OK;
    for (int64_t k = 0; k < n; k++)  // packed predecessor i<<16|p of the -nl kernel
        if (ctx->sides[0].labels(pr[k].contig) > 32767 || ctx->sides[1].labels(pr[k].seg) > 65535)
            return fail(ctx, SA_EINVAL, "-nl (unbanded) supports contigs up to 32767 and segments up to 65535 labels");
    return SA_OK;
}

int n_fill_ctas(const sa_ctx *ctx) { return ctx->n_sms * ctx->ctas_per_sm; }

static std::string cpl_debug_text(const int32_t *e) {
    std::string t;
    if (e[12] == 0 && e[16] == 0) return t;
    t = "; queue head " + std::to_string(e[12]) + "; ring progress (strip:column per unit):";
    for (int k = 0; k < 16 && k < e[3]; k++)
        t += e[16 + k] == 0 ? " unclaimed" : " " + std::to_string((e[16 + k] >> 12) & 0x3ffff) + ":" + std::to_string(e[16 + k] & 0xfff);
    return t;
}

// after a stream synchronisation: did a coupled-unit handshake of the last fill time out (a unit's producer never ran)?
int cpl_check(sa_ctx *ctx) {
    if (!ctx->cpl_pending) return SA_OK;
    ctx->cpl_pending = false;
    const int32_t *e = ctx->h_cpl_err.as<int32_t>();
    if (e[0] != 0)
        return fail(ctx, SA_ECUDA, "fill kernel: a wavefront hand-over between coupled units timed out (problem " + std::to_string(e[1]) + " unit " +
                                       std::to_string(e[2]) + "/" + std::to_string(e[3]) + " strip " + std::to_string(e[4]) + " column " + std::to_string(e[5]) +
                                       ": expected tag " + std::to_string(e[6]) + ", saw " + std::to_string(e[7]) + "; " + std::to_string(e[8]) + " x " +
                                       std::to_string(e[9]) + " labels, queue entry " + std::to_string(e[10]) + ", lane " + std::to_string(e[11]) + ")" + cpl_debug_text(e));
    return SA_OK;
}

int prepare_fill(sa_ctx *ctx, FillParams &P, const sa_mode *mode) {
    if (!mode) return fail(ctx, SA_EINVAL, "mode is NULL");
    if (ctx->sides[0].n_maps == 0 || ctx->sides[1].n_maps == 0) return fail(ctx, SA_EINVAL, "maps not uploaded");
    if (mode->lookback < LB) return fail(ctx, SA_EINVAL, "lookback must be 6 (banded kernels) or larger (-nl, unbanded)");
    memset(&P, 0, sizeof(P));
    P.r_flt_stop_below = -INFINITY;  // rounds: never stop early unless a filter says so
    P.contigs = ctx->sides[0].dev();
    P.segs = ctx->sides[1].dev();
    P.init_mode = mode->init_mode;
    P.start_mode = mode->start_mode;
    P.fitting = mode->fitting;
    fill_powf_constants(P.pk);
    const long long n_warps = (long long)n_fill_ctas(ctx) * FILL_WARPS;
    P.halo_stride = (long long)std::max(ctx->sides[1].max_labels, 1) * 8;
    CU(ctx->d_halo.reserve(sizeof(float) * (size_t)(n_warps * P.halo_stride)));
    CU(ctx->d_halo2.reserve(sizeof(float) * (size_t)((long long)ctx->n_sms * COOP_MINB * P.halo_stride)));
    CU(ctx->d_counter.reserve(8 * sizeof(unsigned long long)));
    if (ctx->n_clusters > 0) {
        CU(ctx->d_halo3.reserve(sizeof(float) * (size_t)((long long)ctx->n_clusters * P.halo_stride)));
        CU(ctx->d_mail.reserve(sizeof(float) * (size_t)ctx->n_clusters * CLUSTER_MAIL_FLOATS));
    }
    return SA_OK;
}

// Launch the fill over an uploaded problem list whose device order array holds n_bulk bulk problems followed
// by n_large large ones.  The two kernels run concurrently (stream / stream2) and join on `stream`.
int launch_fill_classes(sa_ctx *ctx, FillParams P, const int32_t *d_order, const Plan &plan, int store, bool swap) {
    const int64_t n_bulk = plan.n_bulk, n_large = plan.n_large, n_huge = plan.n_huge;
    unsigned long long *counters = ctx->d_counter.as<unsigned long long>();
    CU(cudaMemsetAsync(counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    P.stats = counters + 4;
    int launches = 0;
    const int64_t n_units = (int64_t)plan.units.size();
    if (n_units > 0) {  // coupled units go first (largest problems); they and the bulk kernel drain their queues side by side
        const long long stride = (long long)std::max(plan.cpl_max_nx, 1) * 6;
        CU(ctx->d_units.reserve(sizeof(CplUnit) * (size_t)n_units));
        CU(ctx->d_mbox.reserve(sizeof(unsigned long long) * (size_t)(n_units * stride)));
        CU(ctx->d_cpl_res.reserve(sizeof(CplRes) * (size_t)n_units));
        CU(ctx->d_cpl_done.reserve(sizeof(int32_t) * (size_t)(2 * n_units + 64)));  // + the time-out flag and its description (+ debug progress marks)
        CU(cpy(ctx, ctx->d_units.p, plan.units.data(), sizeof(CplUnit) * (size_t)n_units, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->d_mbox.p, 0, sizeof(unsigned long long) * (size_t)(n_units * stride), ctx->stream));  // tag 0 = empty
        CU(cudaMemsetAsync(ctx->d_cpl_done.p, 0, sizeof(int32_t) * (size_t)(2 * n_units + 64), ctx->stream));
        CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
        FillParams C = P;
        C.order = nullptr;
        C.units = ctx->d_units.as<CplUnit>();
        C.n_problems = n_units;
        C.counter = counters + 1;
        C.mbox = ctx->d_mbox.as<unsigned long long>();
        C.mbox_stride = stride;
        C.cpl_res = ctx->d_cpl_res.as<CplRes>();
        C.cpl_done = ctx->d_cpl_done.as<int32_t>();
        C.cpl_error = ctx->d_cpl_done.as<int32_t>() + n_units;
        ctx->cpl_epoch = (ctx->cpl_epoch % 15u) + 1u;
        C.cpl_epoch = ctx->cpl_epoch << 28;
        // A warp may wait for a unit that no warp has pulled yet (the ring's last unit of a partially claimed problem): at most
        // about one unit's duration (~1.3e7 cells/s per warp).  Ten of those, one poll ~0.2 us, at least 2^25 polls (~6 s).
        C.cpl_spin_limit = (uint32_t)std::min(4.0e9, std::max(33554432.0, 10.0 * plan.cpl_max_unit_cells / 1.3e7 / 2.0e-7));
        int64_t cpl_ctas = std::min<int64_t>((n_units + FILL_WARPS - 1) / FILL_WARPS, n_fill_ctas(ctx));
        if (const char *e = getenv("SA_CPL_MAX_CTAS")) cpl_ctas = std::max<int64_t>(8, std::min<int64_t>(cpl_ctas, atoll(e)));  // (testing aid)
        launch_fill(C, store, swap, 3, (int)cpl_ctas, ctx->stream2);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ev_join, ctx->stream2));
        launches++;
    }
    if (n_large > 0 || n_huge > 0) CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
    if (n_huge > 0) {  // longest first: clusters of CLUSTER_CTAS SMs
        CU(cudaStreamWaitEvent(ctx->stream3, ctx->ev_fork, 0));
        FillParams H = P;
        H.order = d_order + n_bulk + n_large;
        H.n_problems = n_huge;
        H.counter = counters + 2;
        H.halo = ctx->d_halo3.as<float>();
        H.cluster_mail = ctx->d_mail.as<float>();
        const int ncl = (int)std::min<int64_t>(n_huge, ctx->n_clusters);
        launch_fill(H, store, swap, 2, ncl * CLUSTER_CTAS, ctx->stream3);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ev_join3, ctx->stream3));
        launches++;
    }
    if (n_large > 0) {
        CU(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
        FillParams L = P;
        L.order = d_order + n_bulk;
        L.n_problems = n_large;
        L.counter = counters + 1;
        L.halo = ctx->d_halo2.as<float>();
        launch_fill(L, store, swap, 1, (int)std::min<int64_t>(n_large, (int64_t)ctx->n_sms * COOP_MINB), ctx->stream2);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ev_join, ctx->stream2));
        launches++;
    }
    if (n_bulk > 0) {
        P.order = d_order;
        P.n_problems = n_bulk;
        P.counter = counters;
        P.halo = ctx->d_halo.as<float>();
        const int64_t want = (n_bulk + FILL_WARPS - 1) / FILL_WARPS;
        launch_fill(P, store, swap, 0, (int)std::min<int64_t>(want, n_fill_ctas(ctx)), ctx->stream);
        CU(cudaGetLastError());
        launches++;
    }
    if (n_large > 0 || n_units > 0) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    if (n_units > 0) {  // read by cpl_check() after the caller's next synchronisation of the stream
        CU(ctx->h_cpl_err.reserve(64 * sizeof(int32_t)));
        CU(cudaMemcpyAsync(ctx->h_cpl_err.p, ctx->d_cpl_done.as<int32_t>() + n_units, 64 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->cpl_pending = true;
    }
    if (n_huge > 0) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join3, 0));
    ctx->timing.n_launches += launches;
    return SA_OK;
}

}  // namespace

extern "C" {

int sa_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int sa_ctx_create(int device, sa_ctx **out) {
    if (!out) return SA_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return SA_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return SA_ENODEV;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SA_ENODEV;
    if (prop.major != 10) {
        fprintf(stderr, "segaligner_b200: device %d is sm_%d%d; this library contains sm_100a code only\n", device,
                prop.major, prop.minor);
        return SA_ENODEV;
    }
    sa_ctx *ctx = new sa_ctx();
    ctx->device = device;
    ctx->n_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return SA_ECUDA;
    }
    cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->stream3, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_join3, cudaEventDisableTiming);
    ctx->n_clusters = max_clusters();
    for (auto &e : ctx->ev) cudaEventCreate(&e);
    for (auto &e : ctx->ev_user) cudaEventCreate(&e);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    ctx->ctas_per_sm = fill_ctas_per_sm();
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // stored matrices cost 5 bytes per cell; a chunk may take 45 % of free memory (the whole C3 detection sweep, 1.33 x 10^10
    // cells = 67 GB, is then one chunk on a 180 GB B200; the rest holds maps, row maxima, mailboxes and the next allocations)
    ctx->store_budget_cells = (size_t)((double)free_b * 0.45 / 5.0);
    if (const char *e = getenv("SA_STORE_BUDGET_CELLS")) ctx->store_budget_cells = (size_t)atoll(e);
    *out = ctx;
    return SA_OK;
}

void sa_ctx_destroy(sa_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &s : ctx->sides) {
        s.d_pos.release();
        s.d_prob.release();
        s.d_ops.release();
        s.d_off.release();
    }
    DevBuf *bufs[] = {&ctx->d_problems, &ctx->d_order,  &ctx->d_results,  &ctx->d_halo,   &ctx->d_counter, &ctx->d_cell_off,
                      &ctx->d_S,        &ctx->d_code,   &ctx->d_used,     &ctx->d_path_off, &ctx->d_path_j,  &ctx->d_path_q,
                      &ctx->d_path_s,   &ctx->d_tmp0,   &ctx->d_tmp1,     &ctx->d_tmp2,     &ctx->d_halo2,  &ctx->d_row_off, &ctx->d_rowmax_s,
                      &ctx->d_rowmax_q, &ctx->d_prev32, &ctx->d_halo3, &ctx->d_mail, &ctx->d_r_code, &ctx->d_r_halo, &ctx->d_r_col2,
                      &ctx->d_r_lastrow, &ctx->d_r_used, &ctx->d_r_path, &ctx->d_r_ps, &ctx->d_r_thr, &ctx->d_r_rounds, &ctx->d_r_nrounds, &ctx->d_r_roff, &ctx->d_r_packed,
                      &ctx->d_r_pj, &ctx->d_r_pq, &ctx->d_r_psc, &ctx->d_r_ctl,
                      &ctx->d_units, &ctx->d_mbox, &ctx->d_cpl_res, &ctx->d_cpl_done};
    for (auto b : bufs) b->release();
    HostBuf *hbufs[] = {&ctx->h_res, &ctx->h_start, &ctx->h_pj, &ctx->h_pq, &ctx->h_ps, &ctx->h_stage, &ctx->h_cpl_err};
    for (auto b : hbufs) b->release();
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : ctx->ev_user)
        if (e) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_join3) cudaEventDestroy(ctx->ev_join3);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream3) cudaStreamDestroy(ctx->stream3);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *sa_last_error(const sa_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int sa_sync(sa_ctx *ctx) {
    if (!ctx) return SA_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return SA_OK;
}

int sa_set_maps(sa_ctx *ctx, int side, int32_t n_maps, const int64_t *off, const float *pos, const float *prob) {
    if (!ctx) return SA_EINVAL;
    if (side != SA_SIDE_CONTIG && side != SA_SIDE_SEG) return fail(ctx, SA_EINVAL, "bad side");
    if (n_maps <= 0 || !off || !pos || !prob) return fail(ctx, SA_EINVAL, "empty map set");
    CU(cudaSetDevice(ctx->device));
    MapSetHost &S = ctx->sides[side];
    S.off.assign(off, off + n_maps + 1);
    S.n_maps = n_maps;
    S.max_labels = 0;
    const int64_t n_pos = off[n_maps];
    const int64_t n_lab = n_pos - n_maps;
    if (n_pos >= 2147483648ll) return fail(ctx, SA_EINVAL, "map set holds 2^31 or more entries (the kernels index labels with 32 bits)");
    for (int m = 0; m < n_maps; m++) {
        if (off[m + 1] - off[m] < 1) return fail(ctx, SA_EINVAL, "map with no entries");
        S.max_labels = std::max(S.max_labels, S.labels(m));
    }
    std::vector<LabelOps> ops((size_t)std::max<int64_t>(n_lab, 1));
    for (int m = 0; m < n_maps; m++) build_label_ops(pos + off[m], prob + (off[m] - m), S.labels(m), ops.data() + (off[m] - m));
    CU(S.d_pos.reserve(sizeof(float) * (size_t)n_pos));
    CU(S.d_prob.reserve(sizeof(float) * (size_t)std::max<int64_t>(n_lab, 1)));
    CU(S.d_ops.reserve(sizeof(LabelOps) * ops.size()));
    CU(S.d_off.reserve(sizeof(int64_t) * (size_t)(n_maps + 1)));
    CU(cpy(ctx, S.d_pos.p, pos, sizeof(float) * (size_t)n_pos, cudaMemcpyHostToDevice, ctx->stream));
    if (n_lab > 0) {
        CU(cpy(ctx, S.d_prob.p, prob, sizeof(float) * (size_t)n_lab, cudaMemcpyHostToDevice, ctx->stream));
        CU(cpy(ctx, S.d_ops.p, ops.data(), sizeof(LabelOps) * (size_t)n_lab, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cpy(ctx, S.d_off.p, off, sizeof(int64_t) * (size_t)(n_maps + 1), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SA_OK;
}

int sa_plan_create(sa_ctx *ctx, const sa_problem *problems, int64_t n, sa_plan **out) {
    if (!ctx || !out) return SA_EINVAL;
    *out = nullptr;
    if (n <= 0 || !problems) return fail(ctx, SA_EINVAL, "empty problem list");
    CU(cudaSetDevice(ctx->device));
    int rc = check_problems(ctx, problems, n, 0);  // also rejects n >= 2^31
    if (rc) return rc;
    sa_plan *pl = new sa_plan();
    pl->n = n;
    build_plan(ctx, problems, n, n_fill_ctas(ctx) * FILL_WARPS, pl->plan);
    cudaError_t e = pl->d_problems.reserve(sizeof(sa_problem) * (size_t)n);
    if (e == cudaSuccess) e = pl->d_order.reserve(sizeof(int32_t) * (size_t)n);
    if (e == cudaSuccess) e = pl->d_results.reserve(sizeof(sa_result) * (size_t)n);
    if (e == cudaSuccess)
        e = cpy(ctx, pl->d_problems.p, problems, sizeof(sa_problem) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = cpy(ctx, pl->d_order.p, pl->plan.order.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        pl->d_problems.release();
        pl->d_order.release();
        pl->d_results.release();
        delete pl;
        return fail(ctx, e == cudaErrorMemoryAllocation ? SA_ENOMEM : SA_ECUDA, std::string("sa_plan_create: ") + cudaGetErrorString(e));
    }
    *out = pl;
    return SA_OK;
}

void sa_plan_destroy(sa_ctx *ctx, sa_plan *pl) {
    if (!pl) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    pl->d_problems.release();
    pl->d_order.release();
    pl->d_results.release();
    delete pl;
}

int sa_plan_score(sa_ctx *ctx, sa_plan *pl, const sa_mode *mode) {
    if (!ctx || !pl) return SA_EINVAL;
    CU(cudaSetDevice(ctx->device));
    if (mode && mode->lookback != LB) return fail(ctx, SA_EINVAL, "sa_plan_score is banded-only; use sa_score_batch for -nl");
    FillParams P;
    int rc = prepare_fill(ctx, P, mode);
    if (rc) return rc;
    P.problems = pl->d_problems.as<sa_problem>();
    P.results = pl->d_results.as<sa_result>();
    ctx->timing = {-1.0f, 0, 0, 0};  // fill_ms resolved lazily by sa_get_timing
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    rc = launch_fill_classes(ctx, P, pl->d_order.as<int32_t>(), pl->plan, 0, mode->swap_b_x != 0);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    return SA_OK;
}

int sa_plan_results(sa_ctx *ctx, sa_plan *pl, sa_result *results) {
    if (!ctx || !pl || !results) return SA_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cpy(ctx, results, pl->d_results.p, sizeof(sa_result) * (size_t)pl->n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return cpl_check(ctx);
}

int sa_plan_info(const sa_plan *pl, int64_t *n_bulk, int64_t *n_large, double *total_cells) {
    if (!pl) return SA_EINVAL;
    if (n_bulk) *n_bulk = pl->plan.n_bulk;
    if (n_large) *n_large = pl->plan.n_large + pl->plan.n_huge + pl->plan.n_cpl;
    if (total_cells) *total_cells = pl->plan.total_cells;
    return SA_OK;
}

int sa_score_batch(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n, sa_result *results) {
    if (!ctx) return SA_EINVAL;
    if (n <= 0) return SA_OK;
    if (!problems || !results) return fail(ctx, SA_EINVAL, "NULL buffer");
    if (mode && mode->lookback != LB) {  // -nl: the unbanded kernel always stores its matrices; reuse the alignment path
        std::vector<int64_t> poff((size_t)n + 1, 0);
        int32_t dj = 0, dq = 0;
        float ds = 0;
        return sa_align_batch(ctx, mode, problems, n, nullptr, 0, 0, poff.data(), results, &dj, &dq, &ds);
    }
    sa_plan *pl = nullptr;
    int rc = sa_plan_create(ctx, problems, n, &pl);
    if (rc) return rc;
    rc = sa_plan_score(ctx, pl, mode);
    if (rc == SA_OK) rc = sa_plan_results(ctx, pl, results);
    sa_plan_destroy(ctx, pl);
    return rc;
}

// Fill + traceback of n problems; the paths are measured on the device first (count-only traceback + prefix sum) and
// then written PACKED, so only sum(path_len) entries cross PCIe, into pinned host buffers owned by the context.
static int align_packed(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n, const uint32_t *used_bits,
                        int64_t used_words_per_slot, int64_t n_used_slots, int64_t *total_out, bool want_S = false) {
    CU(cudaSetDevice(ctx->device));
    int rc = check_problems(ctx, problems, n, used_bits ? n_used_slots : 0);
    if (rc) return rc;
    rc = check_unbanded_sizes(ctx, mode, problems, n);
    if (rc) return rc;
    FillParams P;
    rc = prepare_fill(ctx, P, mode);
    if (rc) return rc;
    if (used_bits && n_used_slots > 0) {
        if ((double)used_words_per_slot * (double)n_used_slots >= 2147483648.0)
            return fail(ctx, SA_EINVAL, "used-label bitmaps exceed 2^31 words");
        const size_t bytes = sizeof(uint32_t) * (size_t)(used_words_per_slot * n_used_slots);
        CU(ctx->d_used.reserve(bytes));
        CU(cpy(ctx, ctx->d_used.p, used_bits, bytes, cudaMemcpyHostToDevice, ctx->stream));
        P.used_bits = ctx->d_used.as<uint32_t>();
        P.used_words_per_slot = used_words_per_slot;
    }
    ctx->timing = {0, 0, 0, 0};
    CU(ctx->h_res.reserve(sizeof(sa_result) * (size_t)n));
    CU(ctx->h_start.reserve(sizeof(int64_t) * (size_t)(n + 1)));
    // The banded fill stores only the 1-byte predecessor codes (the traceback recomputes the scores along the path
    // exactly); -nl keeps its stored scores and 4-byte predecessors.  `want_S` (the sa_fill_one test hook) keeps S too.
    const bool unbanded_mode = mode->lookback != LB;
    const bool keep_S = unbanded_mode || want_S;
    const size_t budget_cells = keep_S ? ctx->store_budget_cells : ctx->store_budget_cells * 5;
    int64_t total = 0;  // packed path entries so far
    if (n > 100000) {  // big batches: one pinned allocation up front (paths average a handful of entries) instead of regrowing
        const size_t guess = (size_t)n * 8;
        CU(ctx->h_pj.reserve(sizeof(int32_t) * guess));
        CU(ctx->h_pq.reserve(sizeof(int32_t) * guess));
        CU(ctx->h_ps.reserve(sizeof(float) * guess));
    }
    // chunk so that the stored matrices fit the budget
    int64_t k0 = 0;
    std::vector<int64_t> cell_off;
    while (k0 < n) {
        cell_off.clear();
        int64_t cells = 0, k1 = k0;
        while (k1 < n) {
            const int64_t c = (int64_t)ctx->sides[0].labels(problems[k1].contig) * ctx->sides[1].labels(problems[k1].seg);
            if (k1 > k0 && (size_t)(cells + c) > budget_cells) break;
            cell_off.push_back(cells);
            cells += c;
            k1++;
        }
        const int64_t m = k1 - k0;
        const auto t_host0 = std::chrono::steady_clock::now();
        Plan plan;
        build_plan(ctx, problems + k0, m, n_fill_ctas(ctx) * FILL_WARPS, plan);
        const std::vector<int32_t> &order = plan.order;
        CU(ctx->d_problems.reserve(sizeof(sa_problem) * (size_t)m));
        CU(ctx->d_order.reserve(sizeof(int32_t) * (size_t)m));
        CU(ctx->d_results.reserve(sizeof(sa_result) * (size_t)m));
        CU(ctx->d_cell_off.reserve(sizeof(int64_t) * (size_t)m));
        const bool unbanded = mode->lookback != LB;
        if (keep_S) CU(ctx->d_S.reserve(sizeof(float) * (size_t)cells));
        if (unbanded) CU(ctx->d_prev32.reserve(sizeof(int32_t) * (size_t)cells));
        else CU(ctx->d_code.reserve((size_t)cells));
        CU(ctx->d_path_off.reserve(sizeof(int64_t) * (size_t)(m + 1)));
        CU(cpy(ctx, ctx->d_problems.p, problems + k0, sizeof(sa_problem) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        CU(cpy(ctx, ctx->d_order.p, order.data(), sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        CU(cpy(ctx, ctx->d_cell_off.p, cell_off.data(), sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        P.problems = ctx->d_problems.as<sa_problem>();
        P.order = ctx->d_order.as<int32_t>();
        P.cell_off = ctx->d_cell_off.as<int64_t>();
        P.n_problems = m;
        P.results = ctx->d_results.as<sa_result>();
        P.S = keep_S ? ctx->d_S.as<float>() : nullptr;
        P.code = ctx->d_code.as<uint8_t>();
        CU(cudaEventRecord(ctx->ev[0], ctx->stream));
        if (unbanded) {
            P.order = nullptr;
            P.n_problems = m;
            launch_fill_unbanded(P, ctx->d_prev32.as<int32_t>(), mode->swap_b_x != 0,
                                 (int)std::min<int64_t>(m, (int64_t)ctx->n_sms * 8), ctx->stream);
            CU(cudaGetLastError());
            ctx->timing.n_launches += 1;
        } else {
            rc = launch_fill_classes(ctx, P, ctx->d_order.as<int32_t>(), plan, keep_S ? 2 : 1, mode->swap_b_x != 0);
            if (rc) return rc;
        }
        CU(cudaEventRecord(ctx->ev[1], ctx->stream));
        TraceParams T;
        memset(&T, 0, sizeof(T));
        T.prev32 = unbanded ? ctx->d_prev32.as<int32_t>() : nullptr;
        T.contigs = P.contigs;
        T.segs = P.segs;
        T.problems = P.problems;
        T.cell_off = P.cell_off;
        T.n_problems = m;
        T.S = P.S;
        T.code = unbanded ? nullptr : P.code;
        T.results = P.results;
        T.init_mode = mode->init_mode;
        T.swap = mode->swap_b_x != 0;
        // pass 1: measure the paths, prefix-sum the lengths into packed start offsets
        int64_t *d_start = ctx->d_path_off.as<int64_t>();
        T.count_only = 1;
        T.len_out = d_start;
        CU(cudaMemsetAsync(d_start + m, 0, sizeof(int64_t), ctx->stream));
        launch_traceback(T, ctx->stream);
        CU(cudaGetLastError());
        size_t tmp_need = 0;
        exclusive_scan_i64(d_start, m + 1, nullptr, 0, &tmp_need, ctx->stream);
        CU(ctx->d_tmp2.reserve(tmp_need + 16));
        if (exclusive_scan_i64(d_start, m + 1, ctx->d_tmp2.p, ctx->d_tmp2.cap, nullptr, ctx->stream))
            return fail(ctx, SA_ECUDA, "prefix sum of the path lengths failed");
        int64_t n_path = 0;
        CU(cpy(ctx, &n_path, d_start + m, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (int e = cpl_check(ctx)) return e;
        // pass 2: write the paths packed
        CU(ctx->d_path_j.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(n_path, 1)));
        CU(ctx->d_path_q.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(n_path, 1)));
        CU(ctx->d_path_s.reserve(sizeof(float) * (size_t)std::max<int64_t>(n_path, 1)));
        CU(ctx->h_pj.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(total + n_path, 1), true));
        CU(ctx->h_pq.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(total + n_path, 1), true));
        CU(ctx->h_ps.reserve(sizeof(float) * (size_t)std::max<int64_t>(total + n_path, 1), true));
        T.count_only = 0;
        T.len_out = nullptr;
        T.path_off = d_start;
        T.path_j = ctx->d_path_j.as<int32_t>();
        T.path_q = ctx->d_path_q.as<int32_t>();
        T.path_s = ctx->d_path_s.as<float>();
        launch_traceback(T, ctx->stream);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ev[2], ctx->stream));
        CU(cpy(ctx, ctx->h_res.as<sa_result>() + k0, ctx->d_results.p, sizeof(sa_result) * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cpy(ctx, ctx->h_start.as<int64_t>() + k0, d_start, sizeof(int64_t) * (size_t)(m + 1), cudaMemcpyDeviceToHost, ctx->stream));
        if (n_path > 0) {
            CU(cpy(ctx, ctx->h_pj.as<int32_t>() + total, ctx->d_path_j.p, sizeof(int32_t) * (size_t)n_path, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cpy(ctx, ctx->h_pq.as<int32_t>() + total, ctx->d_path_q.p, sizeof(int32_t) * (size_t)n_path, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cpy(ctx, ctx->h_ps.as<float>() + total, ctx->d_path_s.p, sizeof(float) * (size_t)n_path, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CU(cudaStreamSynchronize(ctx->stream));
        if (total > 0) {  // chunk-local start offsets -> global
            int64_t *hs = ctx->h_start.as<int64_t>() + k0;
            for (int64_t k = 0; k <= m; k++) hs[k] += total;
        }
        total += n_path;
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]);
        ctx->timing.fill_ms += a;
        ctx->timing.post_ms += b;
        ctx->timing.n_launches += 2;
        if (getenv("SA_VERBOSE"))
            fprintf(stderr, "  sa_align_batch chunk: %lld problems, %.3e cells, fill %.2f ms (large %lld), traceback %.2f ms, %lld path entries, chunk wall %.1f ms\n",
                    (long long)m, (double)cells, a, (long long)(plan.n_large * 1000 + plan.n_huge + plan.n_cpl), b, (long long)n_path,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count());
        k0 = k1;
    }
    *total_out = total;
    return SA_OK;
}

int sa_align_batch_packed(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n, const uint32_t *used_bits,
                          int64_t used_words_per_slot, int64_t n_used_slots, sa_packed_paths *out) {
    if (!ctx || !out) return SA_EINVAL;
    memset(out, 0, sizeof(*out));
    if (n <= 0) return SA_OK;
    if (!problems || !mode) return fail(ctx, SA_EINVAL, "NULL buffer");
    int64_t total = 0;
    int rc = align_packed(ctx, mode, problems, n, used_bits, used_words_per_slot, n_used_slots, &total);
    if (rc) return rc;
    out->n = n;
    out->total = total;
    out->results = ctx->h_res.as<sa_result>();
    out->start = ctx->h_start.as<int64_t>();
    out->path_j = ctx->h_pj.as<int32_t>();
    out->path_q = ctx->h_pq.as<int32_t>();
    out->path_s = ctx->h_ps.as<float>();
    return SA_OK;
}

int sa_align_batch(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n, const uint32_t *used_bits,
                   int64_t used_words_per_slot, int64_t n_used_slots, const int64_t *path_off, sa_result *results,
                   int32_t *path_j, int32_t *path_q, float *path_s) {
    if (!ctx) return SA_EINVAL;
    if (n <= 0) return SA_OK;
    if (!problems || !results || !path_off || !path_j || !path_q || !path_s || !mode) return fail(ctx, SA_EINVAL, "NULL buffer");
    sa_packed_paths pk;
    int rc = sa_align_batch_packed(ctx, mode, problems, n, used_bits, used_words_per_slot, n_used_slots, &pk);
    if (rc) return rc;
    memcpy(results, pk.results, sizeof(sa_result) * (size_t)n);
    for (int64_t k = 0; k < n; k++) {  // scatter the packed paths to the caller's capacity offsets
        const int64_t cap = path_off[k + 1] - path_off[k];
        int64_t len = pk.results[k].path_len;
        if (len > cap) {  // the caller offered less room than the path needs: truncate like the device loop used to
            len = cap;
            results[k].path_len = (int32_t)cap;
        }
        memcpy(path_j + path_off[k], pk.path_j + pk.start[k], sizeof(int32_t) * (size_t)len);
        memcpy(path_q + path_off[k], pk.path_q + pk.start[k], sizeof(int32_t) * (size_t)len);
        memcpy(path_s + path_off[k], pk.path_s + pk.start[k], sizeof(float) * (size_t)len);
    }
    return SA_OK;
}

// The whole per-pair loop of run_SA_aln on the device (see include/segaligner_b200.h and rounds_finish in sa_kernels.cu).
int sa_align_rounds(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n, const float *thr, int32_t max_alns,
                    const uint32_t *used_bits, int64_t used_words_per_slot, int64_t n_used_slots, sa_rounds_out *out) {
    return sa_align_rounds_filtered(ctx, mode, problems, n, thr, max_alns, used_bits, used_words_per_slot, n_used_slots, nullptr, out);
}

int sa_align_rounds_filtered(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n, const float *thr,
                             int32_t max_alns, const uint32_t *used_bits, int64_t used_words_per_slot, int64_t n_used_slots,
                             const sa_rounds_filter *filter, sa_rounds_out *out) {
    if (!ctx || !out) return SA_EINVAL;
    memset(out, 0, sizeof(*out));
    if (n <= 0) return SA_OK;
    if (!problems || !mode || !thr) return fail(ctx, SA_EINVAL, "NULL buffer");
    if (mode->init_mode != SA_INIT_SEMIGLOBAL || mode->start_mode != SA_START_SEMIGLOBAL || mode->fitting || mode->swap_b_x ||
        mode->lookback != LB)
        return fail(ctx, SA_EINVAL, "sa_align_rounds: semi-global init and start, lookback 6, no swap_b_x, no fitting");
    if (max_alns < 1 || max_alns > 64) return fail(ctx, SA_EINVAL, "sa_align_rounds: max_alns must be 1..64");
    CU(cudaSetDevice(ctx->device));
    int rc = check_problems(ctx, problems, n, used_bits ? n_used_slots : 0);
    if (rc) return rc;
    for (int64_t k = 0; k < n; k++)
        if (ctx->sides[1].labels(problems[k].seg) < 2) return fail(ctx, SA_EINVAL, "sa_align_rounds: segments need >= 2 labels");
    FillParams P;
    rc = prepare_fill(ctx, P, mode);
    if (rc) return rc;
    if (used_bits && n_used_slots > 0) {
        if ((double)used_words_per_slot * (double)n_used_slots >= 2147483648.0)
            return fail(ctx, SA_EINVAL, "used-label bitmaps exceed 2^31 words");
        const size_t bytes = sizeof(uint32_t) * (size_t)(used_words_per_slot * n_used_slots);
        CU(ctx->d_used.reserve(bytes));
        CU(cpy(ctx, ctx->d_used.p, used_bits, bytes, cudaMemcpyHostToDevice, ctx->stream));
        P.used_bits = ctx->d_used.as<uint32_t>();
        P.used_words_per_slot = used_words_per_slot;
    }
    ctx->timing = {0, 0, 0, 0};
    const bool verbose = getenv("SA_VERBOSE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_in = now();
    const int n_ctas = (int)std::min<int64_t>((n + FILL_WARPS - 1) / FILL_WARPS, n_fill_ctas(ctx));
    const long long workers = (long long)n_ctas * FILL_WARPS;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // per-worker scratch ~ 2 bytes per cell of the largest accepted matrix (codes + one 32-byte halo slot per strip column);
    // pairs whose matrix would not fit half of the free memory are left to the caller (n_rounds = -1)
    double scratch_budget = 0.5 * (double)free_b;
    if (const char *e = getenv("SA_ROUNDS_SCRATCH_BYTES")) scratch_budget = atof(e);
    const double cap_cells = scratch_budget / (double)workers / 2.2 - 65536.0;
    std::vector<int64_t> acc;  // accepted problems
    acc.reserve((size_t)n);
    int max_nb = 1, max_nx = 2, max_path = 1;
    long long max_cells = 1, max_halo = 8;
    for (int64_t k = 0; k < n; k++) {
        const long long nb = ctx->sides[0].labels(problems[k].contig), nx = ctx->sides[1].labels(problems[k].seg);
        if ((double)nb * (double)nx > cap_cells) continue;
        acc.push_back(k);
        max_nb = std::max<int>(max_nb, (int)nb);
        max_nx = std::max<int>(max_nx, (int)nx);
        max_path = std::max<int>(max_path, (int)std::min(nb, nx));
        max_cells = std::max(max_cells, nb * nx);
        max_halo = std::max(max_halo, ((nb + 31) / 32) * nx * 8);
    }
    if (filter) {
        if (filter->min_len < 2) return fail(ctx, SA_EINVAL, "sa_align_rounds_filtered: min_len must be >= 2");
        P.r_flt_on = 1;
        P.r_flt_tip = filter->tip ? 1 : 0;
        P.r_flt_min_len = filter->min_len;
        P.r_flt_mean = filter->mean_min;
        P.r_flt_mean_short = filter->mean_min_short;
        if (filter->stop_below > 0.0f) P.r_flt_stop_below = filter->stop_below;
        if (filter->tip && filter->tip_windows) {
            // A path that can be written has n >= min_len entries and score / (n - 1) >= m := 0.999 * min(mean_min, mean_min_short),
            // i.e. penalties of at most c := 10000 - m per step on average.  Swept with a gain of G per step instead of 10000 it
            // is worth >= (G - c)(n - 1) >= (G - c)(min_len - 1) =: W, and so is the cell it ends in (fp32 rounding of the two
            // chains: < 2.1 each, partial sums stay below 2^19; the test value leaves 10).  G trades two kinds of chance paths
            // off against each other: long ones with mediocre steps sink below zero as G falls towards c, short ones with
            // near-perfect steps reach W as soon as (min_len - 2) G >= W.  G = 1.1 (min_len - 1) c keeps the latter just out.
            const float m = 0.999f * std::min(filter->mean_min, filter->mean_min_short);
            if (!(m > 0.0f && m < 10000.0f)) return fail(ctx, SA_EINVAL, "sa_align_rounds_filtered: tip_windows needs 0 < mean_min, mean_min_short < 10000");
            const float c = 10000.0f - m;
            P.r_flt_tipwin = 1;
            P.r_win_gain = std::min(10000.0f, 1.1f * (float)(filter->min_len - 1) * c);
            P.r_win_min = (P.r_win_gain - c) * (float)(filter->min_len - 1) - 10.0f;
            // rows such a path can span per step: itself plus the labels it can afford to skip at 5000 each (fn_t); 0.1 % up
            // for the rounding slack of the chain (2.1 in c (n - 1))
            P.r_win_slope = (1.0f + c / 5000.0f) * 1.001f;
        }
    }
    // Where the caller reads the records: normally straight from the pinned staging buffer (every pair accepted, one chunk: the
    // common case, and at 10^7 pairs the scatter target alone would be gigabytes of page faults); otherwise scattered to
    // per-problem slots in pageable memory, set up when the first chunk shows that it is needed.
    int32_t *h_nr = nullptr;
    const int32_t *direct_nr = nullptr;
    const sa_round *direct_rounds = nullptr;
    const double t_acc = now();
    double t_reserved = t_acc, t_plan = t_acc;
    int64_t total = 0;
    if (!acc.empty()) {
        const int used_words = (max_nb + 31) / 32;
        CU(ctx->d_r_code.reserve((size_t)workers * (size_t)max_cells));
        CU(ctx->d_r_halo.reserve(sizeof(float) * (size_t)workers * (size_t)max_halo));
        CU(ctx->d_r_col2.reserve(sizeof(float) * (size_t)workers * 2 * (size_t)max_nb));
        CU(ctx->d_r_lastrow.reserve(sizeof(float) * (size_t)workers * (size_t)max_nx));
        CU(ctx->d_r_used.reserve(sizeof(uint32_t) * (size_t)workers * (size_t)used_words));
        CU(ctx->d_r_path.reserve(sizeof(int32_t) * (size_t)workers * 2 * (size_t)max_path));
        CU(ctx->d_r_ps.reserve(sizeof(float) * (size_t)workers * (size_t)max_path));
        CU(ctx->d_r_ctl.reserve(64));
        P.r_code = ctx->d_r_code.as<uint8_t>();
        P.r_halo = ctx->d_r_halo.as<float>();
        P.r_col2 = ctx->d_r_col2.as<float>();
        P.r_lastrow = ctx->d_r_lastrow.as<float>();
        P.r_used = ctx->d_r_used.as<uint32_t>();
        P.r_path = ctx->d_r_path.as<int32_t>();
        P.r_ps = ctx->d_r_ps.as<float>();
        P.r_cells = max_cells;
        P.r_halo_floats = max_halo;
        P.r_max_nb = max_nb;
        P.r_max_nx = max_nx;
        P.r_used_words = used_words;
        P.r_max_path = max_path;
        P.r_max_alns = max_alns;
        // the packed path arena is sized for the worst case of a chunk (every round of every pair a full-length path)
        cudaMemGetInfo(&free_b, &total_b);
        double arena_budget = 0.25 * (double)free_b;
        if (const char *e = getenv("SA_ROUNDS_ARENA_BYTES")) arena_budget = atof(e);
        const long long arena_cap = std::max<long long>((long long)(arena_budget / 12.0), (long long)max_alns * max_path);
        std::vector<sa_problem> cp;
        std::vector<float> cthr;
        size_t k0 = 0;
        while (k0 < acc.size()) {
            long long worst = 0;
            size_t k1 = k0;
            cp.clear();
            cthr.clear();
            while (k1 < acc.size()) {
                const sa_problem &pr = problems[acc[k1]];
                const long long w = (long long)max_alns * std::min(ctx->sides[0].labels(pr.contig), ctx->sides[1].labels(pr.seg));
                if (k1 > k0 && worst + w > arena_cap) break;
                worst += w;
                cp.push_back(pr);
                cthr.push_back(thr[acc[k1]]);
                k1++;
            }
            const int64_t m = (int64_t)(k1 - k0);
            const bool direct = k0 == 0 && k1 == acc.size() && (int64_t)acc.size() == n;
            if (!direct && !h_nr) {
                ctx->h_nr.assign((size_t)n, -1);
                ctx->h_rounds.clear();
                h_nr = ctx->h_nr.data();
            }
            t_reserved = now();
            Plan plan;
            build_plan(ctx, cp.data(), m, n_fill_ctas(ctx) * FILL_WARPS, plan);
            // one queue, largest first: the size classes of the plan in descending order (huge, large, bulk)
            const std::vector<int32_t> order = plan.big_first();
            t_plan = now();
            CU(ctx->d_problems.reserve(sizeof(sa_problem) * (size_t)m));
            CU(ctx->d_order.reserve(sizeof(int32_t) * (size_t)m));
            CU(ctx->d_r_thr.reserve(sizeof(float) * (size_t)m));
            CU(ctx->d_r_nrounds.reserve(sizeof(int32_t) * (size_t)m));
            CU(ctx->d_r_rounds.reserve(sizeof(sa_round) * (size_t)m * (size_t)max_alns));
            CU(ctx->d_r_pj.reserve(sizeof(int32_t) * (size_t)std::min(worst, arena_cap)));
            CU(ctx->d_r_pq.reserve(sizeof(int32_t) * (size_t)std::min(worst, arena_cap)));
            CU(ctx->d_r_psc.reserve(sizeof(float) * (size_t)std::min(worst, arena_cap)));
            CU(cpy(ctx, ctx->d_problems.p, cp.data(), sizeof(sa_problem) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
            CU(cpy(ctx, ctx->d_order.p, order.data(), sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
            CU(cpy(ctx, ctx->d_r_thr.p, cthr.data(), sizeof(float) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemsetAsync(ctx->d_r_ctl.p, 0, 64, ctx->stream));
            CU(cudaMemsetAsync(ctx->d_counter.p, 0, 8 * sizeof(unsigned long long), ctx->stream));
            P.problems = ctx->d_problems.as<sa_problem>();
            P.order = ctx->d_order.as<int32_t>();
            P.n_problems = m;
            P.counter = ctx->d_counter.as<unsigned long long>();
            P.stats = ctx->d_counter.as<unsigned long long>() + 4;
            P.r_thr = ctx->d_r_thr.as<float>();
            P.r_rounds = ctx->d_r_rounds.as<sa_round>();
            P.r_n_rounds = ctx->d_r_nrounds.as<int32_t>();
            P.r_arena_top = ctx->d_r_ctl.as<unsigned long long>();
            P.r_overflow = reinterpret_cast<int32_t *>(ctx->d_r_ctl.as<unsigned long long>() + 1);
            P.r_arena_cap = std::min(worst, arena_cap);
            P.r_pj = ctx->d_r_pj.as<int32_t>();
            P.r_pq = ctx->d_r_pq.as<int32_t>();
            P.r_psc = ctx->d_r_psc.as<float>();
            const double t_launch = now();
            CU(cudaEventRecord(ctx->ev[0], ctx->stream));
            launch_fill(P, 3, false, 0, (int)std::min<int64_t>((m + FILL_WARPS - 1) / FILL_WARPS, n_ctas), ctx->stream);
            CU(cudaGetLastError());
            CU(cudaEventRecord(ctx->ev[1], ctx->stream));
            // The round records leave the device PACKED (pair i: n_rounds[i] of them, in pair order): a count + prefix sum + gather
            // behind the fill, then [ctl (16 B)][n_records (8 B)][n_rounds (m)][records] through a pinned staging buffer.
            CU(ctx->d_r_roff.reserve(sizeof(int64_t) * (size_t)(m + 1)));
            int64_t *d_roff = ctx->d_r_roff.as<int64_t>();
            launch_rounds_count(P.r_n_rounds, d_roff, m, ctx->stream);
            CU(cudaGetLastError());
            {
                size_t tmp_need = 0;
                exclusive_scan_i64(d_roff, m + 1, nullptr, 0, &tmp_need, ctx->stream);
                CU(ctx->d_tmp2.reserve(tmp_need + 16));
                if (exclusive_scan_i64(d_roff, m + 1, ctx->d_tmp2.p, ctx->d_tmp2.cap, nullptr, ctx->stream))
                    return fail(ctx, SA_ECUDA, "prefix sum of the round counts failed");
            }
            const size_t nrec_at = 16, nr_at = 32, rr_at = (nr_at + sizeof(int32_t) * (size_t)m + 15) & ~(size_t)15;
            {   // pinned allocations cost 0.3 - 1 s per GB: make them while the kernel runs, from a guess of the volumes (most
                // pairs of a large batch end after a round or two; with a filter only the tracebacks that can become files
                // come back: a few entries per pair at most)
                const int64_t rec_guess = std::min<int64_t>(m * (int64_t)max_alns, m + (1 << 20));
                CU(ctx->h_stage.reserve(rr_at + sizeof(sa_round) * (size_t)rec_guess));
                const int64_t guess = total + std::min<int64_t>((int64_t)std::min(worst, arena_cap), filter ? m * 4 + (1 << 20) : m * (int64_t)max_alns * 14);
                CU(ctx->h_pj.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(guess, 1), true));
                CU(ctx->h_pq.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(guess, 1), true));
                CU(ctx->h_ps.reserve(sizeof(float) * (size_t)std::max<int64_t>(guess, 1), true));
            }
            const double t_pinned = now();
            {
                char *stage = ctx->h_stage.as<char>();
                CU(cpy(ctx, stage, ctx->d_r_ctl.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
                CU(cpy(ctx, stage + nrec_at, d_roff + m, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
                CU(cpy(ctx, stage + nr_at, ctx->d_r_nrounds.p, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
                CU(cudaStreamSynchronize(ctx->stream));
            }
            const int64_t n_rec = *reinterpret_cast<const int64_t *>(ctx->h_stage.as<char>() + nrec_at);
            if ((int32_t)reinterpret_cast<const unsigned long long *>(ctx->h_stage.as<char>())[1] != 0)
                return fail(ctx, SA_ECUDA, "sa_align_rounds: path arena overflow (internal sizing error)");
            const int64_t n_path = (int64_t)reinterpret_cast<const unsigned long long *>(ctx->h_stage.as<char>())[0];
            CU(ctx->d_r_packed.reserve(sizeof(sa_round) * (size_t)std::max<int64_t>(n_rec, 1)));
            launch_rounds_gather(P.r_rounds, P.r_n_rounds, d_roff, ctx->d_r_packed.as<sa_round>(), m, max_alns, ctx->stream);
            CU(cudaGetLastError());
            CU(ctx->h_stage.reserve(rr_at + sizeof(sa_round) * (size_t)std::max<int64_t>(n_rec, 1), true));  // (keeps ctl and n_rounds)
            char *stage = ctx->h_stage.as<char>();
            const int32_t *nr = reinterpret_cast<const int32_t *>(stage + nr_at);
            const sa_round *rr = reinterpret_cast<const sa_round *>(stage + rr_at);
            if (n_rec > 0) CU(cpy(ctx, stage + rr_at, ctx->d_r_packed.p, sizeof(sa_round) * (size_t)n_rec, cudaMemcpyDeviceToHost, ctx->stream));
            const double t_records = now();
            CU(ctx->h_pj.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(total + n_path, 1), true));
            CU(ctx->h_pq.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(total + n_path, 1), true));
            CU(ctx->h_ps.reserve(sizeof(float) * (size_t)std::max<int64_t>(total + n_path, 1), true));
            if (n_path > 0) {
                CU(cpy(ctx, ctx->h_pj.as<int32_t>() + total, ctx->d_r_pj.p, sizeof(int32_t) * (size_t)n_path, cudaMemcpyDeviceToHost, ctx->stream));
                CU(cpy(ctx, ctx->h_pq.as<int32_t>() + total, ctx->d_r_pq.p, sizeof(int32_t) * (size_t)n_path, cudaMemcpyDeviceToHost, ctx->stream));
                CU(cpy(ctx, ctx->h_ps.as<float>() + total, ctx->d_r_psc.p, sizeof(float) * (size_t)n_path, cudaMemcpyDeviceToHost, ctx->stream));
            }
            CU(cudaStreamSynchronize(ctx->stream));
            const double t_paths = now();
            if (direct) {
                direct_nr = nr;
                direct_rounds = rr;
            } else {  // append the chunk's records (pair order = problem order), path offsets made global; n_rounds to their problems
                const size_t at = ctx->h_rounds.size();
                ctx->h_rounds.resize(at + (size_t)n_rec);
                for (int64_t r = 0; r < n_rec; r++) {
                    sa_round v = rr[(size_t)r];
                    if (v.path_off >= 0) v.path_off += total;
                    ctx->h_rounds[at + (size_t)r] = v;
                }
                for (int64_t i = 0; i < m; i++) h_nr[acc[k0 + (size_t)i]] = nr[(size_t)i];
            }
            total += n_path;
            float a = 0;
            cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
            ctx->timing.fill_ms += a;
            ctx->timing.n_launches += 1;
            if (verbose) {
                fprintf(stderr, "  sa_align_rounds chunk: %lld pairs, kernel %.2f ms, %lld path entries, %d workers x %.1f MB scratch\n",
                        (long long)m, a, (long long)n_path, (int)workers,
                        ((double)max_cells + 4.0 * (double)max_halo + 8.0 * max_nb + 4.0 * max_nx + 4.0 * used_words + 12.0 * max_path) / 1e6);
                fprintf(stderr, "    host side (ms since the call): pairs checked at %.1f, scratch reserved at %.1f, plan built at %.1f, launch at %.1f, pinned buffers ready at %.1f, records on the host at %.1f, paths at %.1f, %s at %.1f\n",
                        t_acc - t_in, t_reserved - t_in, t_plan - t_in, t_launch - t_in, t_pinned - t_in, t_records - t_in, t_paths - t_in,
                        direct ? "read in place" : "scattered", now() - t_in);
            }
            k0 = k1;
        }
    }
    out->n = n;
    out->max_alns = max_alns;
    out->total = total;
    if (!direct_nr && !h_nr) {  // nothing was accepted
        ctx->h_nr.assign((size_t)n, -1);
        ctx->h_rounds.clear();
        h_nr = ctx->h_nr.data();
    }
    out->n_rounds = direct_nr ? direct_nr : h_nr;
    out->rounds = direct_nr ? direct_rounds : ctx->h_rounds.data();
    ctx->h_roff.resize((size_t)n + 1);  // where each pair's records start: the prefix sum of the round counts
    {
        int64_t at = 0;
        for (int64_t k = 0; k < n; k++) {
            ctx->h_roff[(size_t)k] = at;
            at += out->n_rounds[k] > 0 ? out->n_rounds[k] : 0;
        }
        ctx->h_roff[(size_t)n] = at;
    }
    out->round_off = ctx->h_roff.data();
    out->path_j = ctx->h_pj.as<int32_t>();
    out->path_q = ctx->h_pq.as<int32_t>();
    out->path_s = ctx->h_ps.as<float>();
    return SA_OK;
}

// Pinned host buffers of later sa_align_rounds* calls, reserved ahead of time (same sizes as the guesses made there).  Touches
// only h_stage / h_pj / h_pq / h_ps, which no scoring call uses, and reports errors by return value only (ctx->err belongs to
// the thread that drives the context).
int sa_prepare_rounds(sa_ctx *ctx, int64_t n_pairs, int32_t max_alns) {
    if (!ctx || n_pairs < 0 || max_alns < 1) return SA_EINVAL;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ECUDA;
    const size_t rr_at = (32 + sizeof(int32_t) * (size_t)n_pairs + 15) & ~(size_t)15;
    const int64_t rec_guess = std::min<int64_t>(n_pairs * (int64_t)max_alns, n_pairs + (1 << 20));
    const int64_t path_guess = n_pairs * 4 + (1 << 20);
    if (ctx->h_stage.reserve(rr_at + sizeof(sa_round) * (size_t)rec_guess) != cudaSuccess ||
        ctx->h_pj.reserve(sizeof(int32_t) * (size_t)path_guess, true) != cudaSuccess ||
        ctx->h_pq.reserve(sizeof(int32_t) * (size_t)path_guess, true) != cudaSuccess ||
        ctx->h_ps.reserve(sizeof(float) * (size_t)path_guess, true) != cudaSuccess) {
        (void)cudaGetLastError();
        return SA_ENOMEM;
    }
    return SA_OK;
}

int sa_fill_one(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problem, const uint32_t *used_bits,
                int64_t used_words, float *S, int32_t *prev, sa_result *result) {
    if (!ctx) return SA_EINVAL;
    if (!problem || !result) return fail(ctx, SA_EINVAL, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    sa_problem pr = *problem;
    int rc = check_problems(ctx, &pr, 1, used_bits ? 1 : 0);
    if (rc) return rc;
    const int nb = ctx->sides[0].labels(pr.contig), nx = ctx->sides[1].labels(pr.seg);
    if (used_bits) pr.used_slot = 0;
    int64_t total = 0;
    rc = align_packed(ctx, mode, &pr, 1, used_bits, used_words, used_bits ? 1 : 0, &total, /*want_S=*/true);
    if (rc) return rc;
    *result = ctx->h_res.as<sa_result>()[0];
    const size_t cells = (size_t)nb * nx;
    if (S) CU(cudaMemcpy(S, ctx->d_S.p, sizeof(float) * cells, cudaMemcpyDeviceToHost));
    if (prev && mode->lookback != LB) {
        std::vector<int32_t> pv(cells);
        CU(cudaMemcpy(pv.data(), ctx->d_prev32.p, sizeof(int32_t) * cells, cudaMemcpyDeviceToHost));
        for (size_t c = 0; c < cells; c++) {
            prev[2 * c] = pv[c] < 0 ? -1 : (pv[c] >> 16);
            prev[2 * c + 1] = pv[c] < 0 ? -1 : (pv[c] & 0xffff);
        }
    } else if (prev) {
        std::vector<uint8_t> code(cells);
        CU(cudaMemcpy(code.data(), ctx->d_code.p, cells, cudaMemcpyDeviceToHost));
        for (int j = 0; j < nb; j++)
            for (int q = 0; q < nx; q++) {
                const size_t c = (size_t)j * nx + q;
                if (code[c] == 0) {
                    prev[2 * c] = -1;
                    prev[2 * c + 1] = -1;
                } else {
                    prev[2 * c] = j - ((code[c] - 1) / LB + 1);
                    prev[2 * c + 1] = q - ((code[c] - 1) % LB + 1);
                }
            }
    }
    return SA_OK;
}

int sa_detect_batch(sa_ctx *ctx, const sa_mode *mode, const sa_problem *problems, int64_t n, const int32_t *n_want,
                    const int64_t *scores_off, float *scores_out, int32_t *n_got) {
    if (!ctx) return SA_EINVAL;
    if (n <= 0) return SA_OK;
    if (!problems || !n_want || !scores_off || !scores_out || !n_got) return fail(ctx, SA_EINVAL, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    int rc = check_problems(ctx, problems, n, 0);
    if (rc) return rc;
    for (int64_t k = 0; k < n; k++)
        if (n_want[k] < 0 || scores_off[k + 1] - scores_off[k] < n_want[k])
            return fail(ctx, SA_EINVAL, "scores_off leaves less room than n_want");
    rc = check_unbanded_sizes(ctx, mode, problems, n);
    if (rc) return rc;
    FillParams P;
    rc = prepare_fill(ctx, P, mode);
    if (rc) return rc;
    ctx->timing = {0, 0, 0, 0};
    int64_t k0 = 0;
    std::vector<int64_t> cell_off, row_off, soff;
    while (k0 < n) {
        cell_off.clear();
        row_off.clear();
        int64_t cells = 0, rows = 0, k1 = k0;
        while (k1 < n) {
            const int64_t nb = ctx->sides[0].labels(problems[k1].contig);
            const int64_t c = nb * ctx->sides[1].labels(problems[k1].seg);
            if (k1 > k0 && (size_t)(cells + c) > ctx->store_budget_cells) break;
            cell_off.push_back(cells);
            row_off.push_back(rows);
            cells += c;
            rows += nb;
            k1++;
        }
        const int64_t m = k1 - k0;
        Plan plan;
        build_plan(ctx, problems + k0, m, n_fill_ctas(ctx) * FILL_WARPS, plan);
        soff.resize((size_t)m + 1);
        for (int64_t k = 0; k <= m; k++) soff[(size_t)k] = scores_off[k0 + k] - scores_off[k0];
        const int64_t n_sc = std::max<int64_t>(soff[(size_t)m], 1);
        CU(ctx->d_problems.reserve(sizeof(sa_problem) * (size_t)m));
        CU(ctx->d_order.reserve(sizeof(int32_t) * (size_t)m));
        CU(ctx->d_results.reserve(sizeof(sa_result) * (size_t)m));
        CU(ctx->d_cell_off.reserve(sizeof(int64_t) * (size_t)m));
        CU(ctx->d_row_off.reserve(sizeof(int64_t) * (size_t)m));
        const bool unbanded = mode->lookback != LB;
        CU(ctx->d_S.reserve(sizeof(float) * (size_t)cells));
        if (unbanded) CU(ctx->d_prev32.reserve(sizeof(int32_t) * (size_t)cells));
        else CU(ctx->d_code.reserve((size_t)cells));
        CU(ctx->d_rowmax_s.reserve(sizeof(float) * (size_t)rows));
        CU(ctx->d_rowmax_q.reserve(sizeof(int32_t) * (size_t)rows));
        CU(ctx->d_path_off.reserve(sizeof(int64_t) * (size_t)(m + 1)));  // reused: scores_off
        CU(ctx->d_path_s.reserve(sizeof(float) * (size_t)n_sc));           // reused: scores_out
        CU(ctx->d_path_j.reserve(sizeof(int32_t) * (size_t)m));            // reused: n_want
        CU(ctx->d_path_q.reserve(sizeof(int32_t) * (size_t)m));            // reused: n_got
        CU(cpy(ctx, ctx->d_problems.p, problems + k0, sizeof(sa_problem) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        CU(cpy(ctx, ctx->d_order.p, plan.order.data(), sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        CU(cpy(ctx, ctx->d_cell_off.p, cell_off.data(), sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        CU(cpy(ctx, ctx->d_row_off.p, row_off.data(), sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        CU(cpy(ctx, ctx->d_path_off.p, soff.data(), sizeof(int64_t) * (size_t)(m + 1), cudaMemcpyHostToDevice, ctx->stream));
        CU(cpy(ctx, ctx->d_path_j.p, n_want + k0, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        P.problems = ctx->d_problems.as<sa_problem>();
        P.cell_off = ctx->d_cell_off.as<int64_t>();
        P.row_off = ctx->d_row_off.as<int64_t>();
        P.results = ctx->d_results.as<sa_result>();
        P.S = ctx->d_S.as<float>();
        P.code = ctx->d_code.as<uint8_t>();
        P.rowmax_s = ctx->d_rowmax_s.as<float>();
        P.rowmax_q = ctx->d_rowmax_q.as<int32_t>();
        CU(cudaEventRecord(ctx->ev[0], ctx->stream));
        if (unbanded) {
            P.order = nullptr;
            P.n_problems = m;
            launch_fill_unbanded(P, ctx->d_prev32.as<int32_t>(), mode->swap_b_x != 0,
                                 (int)std::min<int64_t>(m, (int64_t)ctx->n_sms * 8), ctx->stream);
            CU(cudaGetLastError());
            ctx->timing.n_launches += 1;
        } else {
            rc = launch_fill_classes(ctx, P, ctx->d_order.as<int32_t>(), plan, 2, mode->swap_b_x != 0);
            if (rc) return rc;
        }
        CU(cudaEventRecord(ctx->ev[1], ctx->stream));
        DetectParams D;
        memset(&D, 0, sizeof(D));
        D.prev32 = unbanded ? ctx->d_prev32.as<int32_t>() : nullptr;
        D.contigs = P.contigs;
        D.segs = P.segs;
        D.problems = P.problems;
        D.cell_off = P.cell_off;
        D.row_off = P.row_off;
        D.scores_off = ctx->d_path_off.as<int64_t>();
        D.n_problems = m;
        D.S = P.S;
        D.code = unbanded ? nullptr : P.code;
        D.rowmax_s = P.rowmax_s;
        D.rowmax_q = P.rowmax_q;
        D.n_want = ctx->d_path_j.as<int32_t>();
        D.scores_out = ctx->d_path_s.as<float>();
        D.n_got = ctx->d_path_q.as<int32_t>();
        launch_detect(D, ctx->stream);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ev[2], ctx->stream));
        if (soff[(size_t)m] > 0)
            CU(cpy(ctx, scores_out + scores_off[k0], ctx->d_path_s.p, sizeof(float) * (size_t)soff[(size_t)m], cudaMemcpyDeviceToHost, ctx->stream));
        CU(cpy(ctx, n_got + k0, ctx->d_path_q.p, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (int e = cpl_check(ctx)) return e;
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]);
        ctx->timing.fill_ms += a;
        ctx->timing.post_ms += b;
        ctx->timing.n_launches += 1;
        if (getenv("SA_VERBOSE"))
            fprintf(stderr, "  sa_detect_batch chunk: %lld problems, %.3e cells, fill %.2f ms (large %lld), harvest %.2f ms\n",
                    (long long)m, (double)cells, a, (long long)(plan.n_large * 1000 + plan.n_huge + plan.n_cpl), b);
        k0 = k1;
    }
    return SA_OK;
}

int sa_powf12(sa_ctx *ctx, const float *x, int64_t n, float *out, int use_fast_path) {
    if (!ctx) return SA_EINVAL;
    if (n <= 0) return SA_OK;
    CU(cudaSetDevice(ctx->device));
    CU(ctx->d_tmp0.reserve(sizeof(float) * (size_t)n));
    CU(ctx->d_tmp1.reserve(sizeof(float) * (size_t)n));
    CU(cpy(ctx, ctx->d_tmp0.p, x, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    launch_powf12(ctx->d_tmp0.as<float>(), n, ctx->d_tmp1.as<float>(), use_fast_path, ctx->stream);
    CU(cudaGetLastError());
    CU(cpy(ctx, out, ctx->d_tmp1.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SA_OK;
}

int sa_powf12_range(sa_ctx *ctx, uint32_t lo_bits, uint32_t hi_bits, float *out_fast, float *out_generic) {
    if (!ctx) return SA_EINVAL;
    if (hi_bits <= lo_bits) return SA_OK;
    const size_t n = (size_t)hi_bits - lo_bits;
    CU(cudaSetDevice(ctx->device));
    CU(ctx->d_tmp0.reserve(sizeof(float) * n));
    CU(ctx->d_tmp1.reserve(sizeof(float) * n));
    launch_powf12_range(lo_bits, hi_bits, ctx->d_tmp0.as<float>(), ctx->d_tmp1.as<float>(), ctx->stream);
    CU(cudaGetLastError());
    CU(cpy(ctx, out_fast, ctx->d_tmp0.p, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cpy(ctx, out_generic, ctx->d_tmp1.p, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SA_OK;
}

int sa_check_prune_bound(sa_ctx *ctx, uint32_t lo_bits, uint32_t hi_bits, uint64_t *n_violations, float *min_ratio) {
    if (!ctx || !n_violations || !min_ratio) return SA_EINVAL;
    *n_violations = 0;
    *min_ratio = 2.0f;
    if (hi_bits <= lo_bits) return SA_OK;
    CU(cudaSetDevice(ctx->device));
    CU(ctx->d_tmp0.reserve(16));
    unsigned long long *d_n = ctx->d_tmp0.as<unsigned long long>();
    float *d_r = reinterpret_cast<float *>(d_n + 1);
    const float two = 2.0f;
    CU(cudaMemsetAsync(d_n, 0, 8, ctx->stream));
    CU(cpy(ctx, d_r, &two, 4, cudaMemcpyHostToDevice, ctx->stream));
    launch_prune_bound(lo_bits, hi_bits, d_n, d_r, ctx->stream);
    CU(cudaGetLastError());
    unsigned long long h_n = 0;
    CU(cpy(ctx, &h_n, d_n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cpy(ctx, min_ratio, d_r, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *n_violations = h_n;
    return SA_OK;
}

int sa_get_prune_stats(sa_ctx *ctx, uint64_t *n_replays, uint64_t *n_rescans) {
    if (!ctx || !n_replays || !n_rescans) return SA_EINVAL;
    *n_replays = *n_rescans = 0;
    if (!ctx->d_counter.p) return SA_OK;
    CU(cudaSetDevice(ctx->device));
    unsigned long long h[2] = {0, 0};
    CU(cpy(ctx, h, ctx->d_counter.as<unsigned long long>() + 4, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *n_replays = h[0];
    *n_rescans = h[1];
    return SA_OK;
}

int sa_get_timing(sa_ctx *ctx, sa_timing *t) {
    if (!ctx || !t) return SA_EINVAL;
    if (ctx->timing.fill_ms < 0) {
        CU(cudaSetDevice(ctx->device));
        CU(cudaEventSynchronize(ctx->ev[1]));
        float a = 0;
        CU(cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]));
        ctx->timing.fill_ms = a;
    }
    *t = ctx->timing;
    return SA_OK;
}

int sa_digest(sa_ctx *ctx, const char *seq, int64_t n, const char *motif, int32_t motif_len, int32_t offset, int64_t cap,
              int64_t *pos_out, int64_t *n_out) {
    if (!ctx || !seq || !motif || !n_out || n < 0 || cap < 0 || (cap > 0 && !pos_out)) return ctx ? fail(ctx, SA_EINVAL, "bad argument") : SA_EINVAL;
    CU(cudaSetDevice(ctx->device));
    const size_t padded = ((size_t)n + 31) / 32 * 32 + 128;  // the kernel reads whole 32-byte chunks plus 8 bytes of look-ahead (and prefetches one chunk ahead)
    CU(ctx->d_S.reserve(padded));  // sequence bytes (the DP matrices are not live during a digestion)
    CU(ctx->d_tmp2.reserve(64));
    const int64_t dev_cap = std::max<int64_t>(cap, 1);
    CU(ctx->d_path_s.reserve(sizeof(long long) * (size_t)dev_cap));
    CU(cudaMemsetAsync(reinterpret_cast<char *>(ctx->d_S.p) + ((size_t)n / 16) * 16, 0, padded - ((size_t)n / 16) * 16, ctx->stream));
    if (n > 0) CU(cpy(ctx, ctx->d_S.p, seq, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_tmp2.p, 0, 64, ctx->stream));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    int rc = sa_digest_run(ctx->stream, ctx->n_sms, ctx->d_S.as<uint8_t>(), n, motif, motif_len, offset,
                           ctx->d_path_s.as<long long>(), ctx->d_tmp2.as<unsigned long long>(), dev_cap);
    if (rc) return fail(ctx, rc, "sa_digest: the motif must be 4..8 letters of ACGT");
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    unsigned long long cnt = 0;
    CU(cpy(ctx, &cnt, ctx->d_tmp2.p, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->timing = {ms, 0, 1, 0};
    if ((int64_t)cnt > cap) {  // sizing call (or buffer too small): report the raw match count, an upper bound
        *n_out = (int64_t)cnt;
        return SA_OK;
    }
    std::vector<long long> v((size_t)cnt);
    if (cnt) CU(cudaMemcpy(v.data(), ctx->d_path_s.p, sizeof(long long) * (size_t)cnt, cudaMemcpyDeviceToHost));
    std::sort(v.begin(), v.end());  // sorted(list(set(...)))  (generate_cmap.py:123)
    v.erase(std::unique(v.begin(), v.end()), v.end());
    for (size_t k = 0; k < v.size(); k++) pos_out[k] = v[k];
    *n_out = (int64_t)v.size();
    return SA_OK;
}

int sa_event_record(sa_ctx *ctx, int32_t slot) {
    if (!ctx || slot < 0 || slot > 1) return SA_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev_user[slot], ctx->stream));
    return SA_OK;
}

int sa_event_elapsed_ms(sa_ctx *ctx, float *ms) {
    if (!ctx || !ms) return SA_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->ev_user[1]));
    CU(cudaEventElapsedTime(ms, ctx->ev_user[0], ctx->ev_user[1]));
    return SA_OK;
}

int sa_measure_fp64_peak(sa_ctx *ctx, double *dfma_per_s) {
    if (!ctx || !dfma_per_s) return SA_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(ctx->d_tmp2.reserve(64));
    const int iters = 4096, ctas = ctx->n_sms * 8;
    launch_dfma_peak(ctx->d_tmp2.as<double>(), 64, ctas, ctx->stream);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CU(cudaEventRecord(ctx->ev[0], ctx->stream));
        launch_dfma_peak(ctx->d_tmp2.as<double>(), iters, ctas, ctx->stream);
        CU(cudaEventRecord(ctx->ev[1], ctx->stream));
        CU(cudaEventSynchronize(ctx->ev[1]));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        // per thread: iters * 8 unrolled * 8 chains DFMA; thread-level instruction count
        const double inst = (double)ctas * 256.0 * iters * 64.0;
        best = std::max(best, inst / (ms * 1e-3));
    }
    CU(cudaGetLastError());
    *dfma_per_s = best;
    return SA_OK;
}

int sa_get_traffic(sa_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes) {
    if (!ctx || !h2d_bytes || !d2h_bytes) return SA_EINVAL;
    *h2d_bytes = ctx->h2d_bytes;
    *d2h_bytes = ctx->d2h_bytes;
    return SA_OK;
}

int sa_measure_mufu_peak(sa_ctx *ctx, double *mufu_per_s) {
    if (!ctx || !mufu_per_s) return SA_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(ctx->d_tmp2.reserve(64));
    const int iters = 1024, ctas = ctx->n_sms * 8;
    launch_mufu_peak(ctx->d_tmp2.as<float>(), 16, ctas, ctx->stream);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CU(cudaEventRecord(ctx->ev[0], ctx->stream));
        launch_mufu_peak(ctx->d_tmp2.as<float>(), iters, ctas, ctx->stream);
        CU(cudaEventRecord(ctx->ev[1], ctx->stream));
        CU(cudaEventSynchronize(ctx->ev[1]));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        const double inst = (double)ctas * 256.0 * iters * 64.0;  // per thread: iters x 4 x (8 lg2 + 8 ex2)
        best = std::max(best, inst / (ms * 1e-3));
    }
    CU(cudaGetLastError());
    *mufu_per_s = best;
    return SA_OK;
}

}  // extern "C"
