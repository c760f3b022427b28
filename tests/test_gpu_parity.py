"""GPU parity: the sm_100a kernels (through the C ABI) against the oracle restatement, bit-exact.

Inputs are seeded; sizes are what the oracle finishes in seconds.  Covers the three initialisers, both
score_f role assignments (swap_b_x), used-label masks, fitting mode, multi-strip contigs (nb > 32) and
integer-valued maps where exact score ties occur (tie-breaking order of dp_align, SegAligner.cpp:109-127).
"""
import numpy as np
import pytest

import helpers
from ampliconreconstructorom_b200 import capi

pytestmark = pytest.mark.gpu

MODES = [
    # (init, fitting, swap, start)
    (capi.SA_INIT_SEMIGLOBAL, 0, 0, capi.SA_START_SEMIGLOBAL),   # standard scoring / alignment
    (capi.SA_INIT_SEMIGLOBAL, 0, 1, capi.SA_START_LOCAL),        # detection phase 1
    (capi.SA_INIT_LOCAL, 0, 1, capi.SA_START_LOCAL),             # detection phase 3
    (capi.SA_INIT_LOCAL, 0, 0, capi.SA_START_LOCAL),             # -local
    (capi.SA_INIT_FITTING, 1, 0, capi.SA_START_END),             # -fitting
]


def build_sets(rng, oracle, n_pairs, gen, integer_every=3, max_nb=150, max_nx=70):
    contigs, segs = {}, {}
    for k in range(n_pairs):
        nb = int(rng.integers(2, max_nb))
        nx = int(rng.integers(2, max_nx))
        b, x = helpers.related_maps(rng, nb, nx, integer=(k % integer_every == 0))
        contigs[k + 1] = b
        segs[k + 1] = x
    cp = oracle.probs(contigs, gen)
    sp = oracle.probs(segs, gen)
    return contigs, segs, cp, sp


def upload(ctx, maps, probs, side):
    ids = sorted(maps)
    off = np.zeros(len(ids) + 1, dtype=np.int64)
    for k, i in enumerate(ids):
        off[k + 1] = off[k] + len(maps[i])
    pos = np.concatenate([maps[i] for i in ids]).astype(np.float32)
    prob = np.concatenate([probs[i] for i in ids]).astype(np.float32)
    ctx.set_maps(side, off, pos, prob)
    return ids


@pytest.fixture(params=["bulk", "coupled", "coop", "cluster"])
def kernel_class(request, monkeypatch):
    """Run a parity case through every way a problem can be swept: one warp per problem, coupled units (several warps
    anywhere on the GPU handing strip boundaries over through tagged mailboxes; the default for large problems), and the two
    barrier-synchronised shapes of round 1 (8-warp CTA wavefront, 8-CTA x 16-warp cluster wavefront; SA_SHAPES=legacy).
    Forced by absurdly low size thresholds.  Requested only by the tests whose result depends on the shape."""
    if request.param == "coupled":
        monkeypatch.setenv("SA_SHAPES", "coupled")
        monkeypatch.setenv("SA_CPL_ALPHA", "1e-9")
    else:
        monkeypatch.setenv("SA_SHAPES", "legacy")
        monkeypatch.setenv("SA_COOP_ALPHA", "0" if request.param == "bulk" else "1e-9")
        monkeypatch.setenv("SA_NO_CLUSTER", "0" if request.param == "cluster" else "1")
        monkeypatch.setenv("SA_HUGE_MIN_CELLS", "0")
    return request.param


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("gen", [1, 2])
def test_fill_one_bit_exact(ctx, oracle, gen, kernel_class):
    rng = np.random.default_rng(100 + gen)
    contigs, segs, cp, sp = build_sets(rng, oracle, 24, gen)
    cids = upload(ctx, contigs, cp, capi.SA_SIDE_CONTIG)
    sids = upload(ctx, segs, sp, capi.SA_SIDE_SEG)
    for k in range(len(cids)):
        b, x = contigs[cids[k]], segs[sids[k]]
        for (init, fit, swap, start) in MODES:
            used = None
            if k % 4 == 1 and not fit:
                used = sorted(set(rng.integers(0, b.size - 1, size=4).tolist()))
            S, prev, res = ctx.fill_one(ctx.mode(init, start, fit, 6, swap), k, k, used_rows=used)
            So, prevo = oracle.fill(b, x, cp[cids[k]], sp[sids[k]], init, fit, 6, used, swap)
            assert np.array_equal(bits(S), bits(So)), f"S differs: pair {k} mode {(init, fit, swap)}"
            assert np.array_equal(prev, prevo), f"previous differs: pair {k} mode {(init, fit, swap)}"
            st = oracle.start(So, start)
            assert (int(res["j"]), int(res["q"])) == st
            if st[0] >= 0:
                assert bits(res["score"]) == bits(So[st])


def test_score_and_align_batch(ctx, oracle, kernel_class):
    rng = np.random.default_rng(7)
    gen = 2
    contigs, segs, cp, sp = build_sets(rng, oracle, 40, gen, max_nb=300, max_nx=120)
    cids = upload(ctx, contigs, cp, capi.SA_SIDE_CONTIG)
    sids = upload(ctx, segs, sp, capi.SA_SIDE_SEG)
    # all-vs-all problem list
    cc, ss = np.meshgrid(np.arange(len(cids)), np.arange(len(sids)), indexing="ij")
    pr = ctx.problems(cc.ravel(), ss.ravel())
    for (init, fit, swap, start) in MODES[:4]:
        mode = ctx.mode(init, start, fit, 6, swap)
        res = ctx.score_batch(mode, pr)
        # check a seeded subset against the oracle (all of them would take minutes in pure C + Python glue)
        pick = rng.choice(pr.size, size=120, replace=False)
        ares, poff, pj, pq, ps = ctx.align_batch(mode, pr[pick])
        for n, k in enumerate(pick):
            c, s = int(pr["contig"][k]), int(pr["seg"][k])
            So, prevo = oracle.fill(contigs[cids[c]], segs[sids[s]], cp[cids[c]], sp[sids[s]], init, fit, 6, None, swap)
            st = oracle.start(So, start)
            assert (int(res["j"][k]), int(res["q"][k])) == st
            assert bits(res["score"][k]) == bits(So[st])
            oj, oq, os_ = oracle.traceback(So, prevo, *st)
            L = int(ares["path_len"][n])
            assert L == oj.size
            sl = slice(int(poff[n]), int(poff[n]) + L)
            assert np.array_equal(pj[sl], oj) and np.array_equal(pq[sl], oq)
            assert np.array_equal(bits(ps[sl]), bits(os_))
            assert bits(ares["score"][n]) == bits(So[st])


def test_used_label_masks_in_batch(ctx, oracle, kernel_class):
    rng = np.random.default_rng(11)
    contigs, segs, cp, sp = build_sets(rng, oracle, 12, 1, max_nb=120, max_nx=60)
    cids = upload(ctx, contigs, cp, capi.SA_SIDE_CONTIG)
    sids = upload(ctx, segs, sp, capi.SA_SIDE_SEG)
    n = len(cids)
    words = (max(len(contigs[c]) - 1 for c in cids) + 31) // 32
    used_bits = np.zeros((n, words), dtype=np.uint32)
    used_sets = []
    for k in range(n):
        nb = len(contigs[cids[k]]) - 1
        u = sorted(set(rng.integers(0, nb, size=max(nb // 5, 1)).tolist()))
        used_sets.append(u)
        for j in u:
            used_bits[k, j >> 5] |= np.uint32(1 << (j & 31))
    pr = ctx.problems(np.arange(n), np.arange(n), used_slot=np.arange(n))
    mode = ctx.mode(capi.SA_INIT_SEMIGLOBAL, capi.SA_START_SEMIGLOBAL, 0, 6, 0)
    ares, poff, pj, pq, ps = ctx.align_batch(mode, pr, used_bits.ravel(), words)
    for k in range(n):
        So, prevo = oracle.fill(contigs[cids[k]], segs[sids[k]], cp[cids[k]], sp[sids[k]], 0, 0, 6, used_sets[k], 0)
        st = oracle.start(So, 0)
        assert (int(ares["j"][k]), int(ares["q"][k])) == st
        oj, oq, os_ = oracle.traceback(So, prevo, *st)
        sl = slice(int(poff[k]), int(poff[k]) + int(ares["path_len"][k]))
        assert np.array_equal(pj[sl], oj) and np.array_equal(pq[sl], oq) and np.array_equal(bits(ps[sl]), bits(os_))
        # a used row keeps its initial value but may still START a path (dp_align only skips it as a target)
        assert not set(pj[sl][:-1].tolist()) & set(used_sets[k])


def test_packed_alignment_output_equals_capacity_form(ctx, oracle, kernel_class):
    """sa_align_batch_packed (what the executable uses: tracebacks packed by a device prefix sum, pinned buffers owned by
    the context) against sa_align_batch with caller-provided capacity offsets, with and without used-label bitmaps."""
    rng = np.random.default_rng(23)
    contigs, segs, cp, sp = build_sets(rng, oracle, 16, 2, max_nb=260, max_nx=90)
    cids = upload(ctx, contigs, cp, capi.SA_SIDE_CONTIG)
    sids = upload(ctx, segs, sp, capi.SA_SIDE_SEG)
    cc, ss = np.meshgrid(np.arange(len(cids)), np.arange(len(sids)), indexing="ij")
    pr = ctx.problems(cc.ravel(), ss.ravel())
    words = (max(len(contigs[c]) for c in cids) + 31) // 32
    used = rng.integers(0, 2 ** 32, size=(pr.size, words), dtype=np.uint64).astype(np.uint32) & \
        rng.integers(0, 2 ** 32, size=(pr.size, words), dtype=np.uint64).astype(np.uint32)
    for init, fit, swap, start in MODES[:4]:
        mode = ctx.mode(init, start, fit, 6, swap)
        for ub, wps, slots in ((None, 0, False), (used.ravel(), words, True)):
            p = pr.copy()
            p["used_slot"] = np.arange(pr.size) if slots else -1
            res, poff, pj, pq, ps = ctx.align_batch(mode, p, ub, wps)
            r2, st, j2, q2, s2 = ctx.align_batch_packed(mode, p, ub, wps)
            assert res.tobytes() == r2.tobytes()
            assert int(st[-1]) == int(res["path_len"].sum()) == j2.size
            for k in range(pr.size):
                L = int(res["path_len"][k])
                a, b = int(poff[k]), int(st[k])
                assert np.array_equal(pj[a:a + L], j2[b:b + L]) and np.array_equal(pq[a:a + L], q2[b:b + L])
                assert np.array_equal(bits(ps[a:a + L]), bits(s2[b:b + L]))


def test_powf12_device_matches_host_libm(ctx, oracle):
    """Both device paths of powf(x,1.2f) against the host libm the reference links, on 2^22 random bit patterns
    plus every boundary of the fast-path domain and the special values."""
    rng = np.random.default_rng(5)
    u = rng.integers(0, 0x7f800000, size=1 << 22, dtype=np.uint32)
    edge = np.array([0, 1, 0x007fffff, 0x00800000, 0x35330000 - 1, 0x35330000, 0x3f330000, 0x3f800000,
                     0x4f330000 - 1, 0x4f330000, 0x7f7fffff, 0x7f800000, 0x7fc00000], dtype=np.uint32)
    x = np.concatenate([u, edge]).view(np.float32)
    import ctypes as C
    exp = np.array([oracle.lib.sao_powf12(C.c_float(float(v))) for v in x[-edge.size:]], dtype=np.float32)
    fast = ctx.powf12(x, fast=True)
    slow = ctx.powf12(x, fast=False)
    nanmask = np.isnan(slow)
    assert np.array_equal(bits(fast)[~nanmask], bits(slow)[~nanmask])
    assert np.array_equal(np.isnan(fast), nanmask)
    e_nan = np.isnan(exp)
    assert np.array_equal(bits(slow[-edge.size:])[~e_nan], bits(exp)[~e_nan])
    # bulk comparison with numpy's powf (same libm) is not guaranteed bit-equal; use the oracle on a sample
    samp = rng.choice(u.size, size=20000, replace=False)
    ref = np.array([oracle.lib.sao_powf12(C.c_float(float(v))) for v in x[samp]], dtype=np.float32)
    assert np.array_equal(bits(fast[samp]), bits(ref))


def test_powf12_device_exhaustive(ctx, oracle):
    """EVERY non-negative float bit pattern (zero, subnormals, all normals, +inf, the first NaNs) through both device
    paths of powf(x, 1.2f) -- the table replay with its out-of-domain patch, and the generic path -- against the host libm
    powf the reference links (SegAligner.cpp:88), bit for bit: 2^31 device evaluations per path, compared on the host by
    oracle/sa_oracle.c:sao_powf12_cmp_range on all cores."""
    import ctypes as C
    import os
    from concurrent.futures import ThreadPoolExecutor
    cmp_ = oracle.lib.sao_powf12_cmp_range
    cmp_.restype = C.c_longlong
    cmp_.argtypes = [C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    n_thr = max(1, min(os.cpu_count() or 1, 32))
    step = 1 << 27
    end = 0x7f800010
    bad_total, first = 0, None
    with ThreadPoolExecutor(n_thr) as pool:
        for lo in range(0, end, step):
            hi = min(lo + step, end)
            fast, gen = ctx.powf12_range(lo, hi)
            cuts = np.linspace(lo, hi, n_thr + 1).astype(np.int64)

            def part(t):
                a, b = int(cuts[t]), int(cuts[t + 1])
                fb = C.c_uint(0)
                n = cmp_(a, b, fast[a - lo:].ctypes.data, gen[a - lo:].ctypes.data, C.byref(fb))
                return n, fb.value
            for n, fb in pool.map(part, range(n_thr)):
                if n and first is None:
                    first = fb
                bad_total += n
    assert bad_total == 0, f"{bad_total} bit patterns differ from host powf; first 0x{first:08x}"


def test_prune_bound_holds_for_every_float(ctx):
    """The fill kernel prunes candidates with an fp32 lower bound of powf(d, 1.2f) (two MUFU approximations); the
    pruning is exact only if that bound never exceeds the bit-exact replay.  Checked for EVERY non-negative float
    (zero, subnormals, the whole normal range, +inf) on the device, 2^31 evaluations."""
    total = 0
    worst = 2.0
    step = 1 << 28
    for lo in range(0, 0x7f800001, step):
        hi = min(lo + step, 0x7f800001)
        n, r = ctx.check_prune_bound(lo, hi)
        total += n
        worst = min(worst, r)
    assert total == 0, f"{total} floats where the pruning bound exceeds the exact powf"
    assert worst > 1.0  # strict slack everywhere (the margin PRUNE_C is 2^-14 in the log2 domain)


def test_pruning_statistics_and_tie_rescans(ctx, oracle, kernel_class):
    """The fill kernel replays powf bit-exactly once per cell and rescans a cell exactly when the winner is not
    strictly ahead of every other candidate's upper bound.  Integer-valued maps make exact ties common: the rescan path
    must run (n_rescans > 0), about one replay per cell must be executed, and results must still equal the oracle."""
    rng = np.random.default_rng(77)
    contigs, segs = {}, {}
    for k in range(12):
        b, x = helpers.related_maps(rng, int(rng.integers(60, 200)), int(rng.integers(20, 60)), integer=True)
        contigs[k + 1], segs[k + 1] = b, x
    cp, sp = oracle.probs(contigs, 2), oracle.probs(segs, 2)
    cids = upload(ctx, contigs, cp, capi.SA_SIDE_CONTIG)
    sids = upload(ctx, segs, sp, capi.SA_SIDE_SEG)
    cc, ss = np.meshgrid(np.arange(len(cids)), np.arange(len(sids)), indexing="ij")
    pr = ctx.problems(cc.ravel(), ss.ravel())
    mode = ctx.mode(capi.SA_INIT_SEMIGLOBAL, capi.SA_START_SEMIGLOBAL, 0, 6, 0)
    res = ctx.score_batch(mode, pr)
    replays, rescans = ctx.prune_stats()
    cells = sum((len(contigs[cids[int(c)]]) - 1) * (len(segs[sids[int(s)]]) - 1) for c, s in zip(pr["contig"], pr["seg"]))
    # lanes of a partly filled 32-row strip replay too (clamped to the last row), hence the padded upper limit
    padded = sum(-(-(len(contigs[cids[int(c)]]) - 1) // 32) * 32 * (len(segs[sids[int(s)]]) - 1)
                 for c, s in zip(pr["contig"], pr["seg"]))
    assert 0.5 * cells < replays <= padded, (replays, cells, padded)
    assert 0 < rescans < 0.2 * cells, (rescans, cells)
    for k in range(0, pr.size, 7):
        c, s = int(pr["contig"][k]), int(pr["seg"][k])
        So, _ = oracle.fill(contigs[cids[c]], segs[sids[s]], cp[cids[c]], sp[sids[s]], 0, 0, 6, None, 0)
        st = oracle.start(So, 0)
        assert (int(res["j"][k]), int(res["q"][k])) == st and bits(res["score"][k]) == bits(So[st])


def test_unbanded_lookback_nl(ctx, oracle):
    """-nl: lookback = 9999999 (main.cpp:475-478) through the generic kernel, all modes."""
    rng = np.random.default_rng(31)
    contigs, segs, cp, sp = build_sets(rng, oracle, 6, 2, max_nb=45, max_nx=28)
    cids = upload(ctx, contigs, cp, capi.SA_SIDE_CONTIG)
    sids = upload(ctx, segs, sp, capi.SA_SIDE_SEG)
    for k in range(len(cids)):
        for (init, fit, swap, start) in MODES:
            used = [1, 4] if (k == 2 and not fit and len(contigs[cids[k]]) > 6) else None
            S, prev, res = ctx.fill_one(ctx.mode(init, start, fit, 9999999, swap), k, k, used_rows=used)
            So, prevo = oracle.fill(contigs[cids[k]], segs[sids[k]], cp[cids[k]], sp[sids[k]], init, fit, 9999999, used, swap)
            assert np.array_equal(bits(S), bits(So)) and np.array_equal(prev, prevo)
            assert (int(res["j"]), int(res["q"])) == oracle.start(So, start)
    pr = ctx.problems(np.arange(len(cids)), np.arange(len(cids)))
    mode = ctx.mode(0, 0, 0, 9999999, 0)
    res = ctx.score_batch(mode, pr)
    ares, poff, pj, pq, ps = ctx.align_batch(mode, pr)
    for k in range(pr.size):
        So, prevo = oracle.fill(contigs[cids[k]], segs[sids[k]], cp[cids[k]], sp[sids[k]], 0, 0, 9999999, None, 0)
        st = oracle.start(So, 0)
        assert (int(res["j"][k]), int(res["q"][k])) == st and bits(res["score"][k]) == bits(So[st])
        oj, oq, os_ = oracle.traceback(So, prevo, *st)
        sl = slice(int(poff[k]), int(poff[k]) + int(ares["path_len"][k]))
        assert np.array_equal(pj[sl], oj) and np.array_equal(pq[sl], oq) and np.array_equal(bits(ps[sl]), bits(os_))
