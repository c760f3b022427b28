"""GPU parity of sa_align_rounds -- the whole per-pair loop of run_SA_aln (/root/reference/SegAligner/main.cpp:262-300) on the
device, later rounds re-sweeping only the rows the previous path can influence -- against the oracle run round by round with
FULL matrices: every round's backtrack start, score bits, traceback and scores along it must be identical, and so must the
number of rounds (acceptance rule `best > exp_thresh`, compute_partial_score_threshold in fp32, max_alns cap)."""
import numpy as np
import pytest

import helpers
from ampliconreconstructorom_b200 import capi, synth

pytestmark = pytest.mark.gpu
f32 = np.float32


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def reference_rounds(oracle, b, x, bp, xp, thr, max_alns, seed):
    """run_SA_aln's while loop (main.cpp:268-300) with the oracle's full-matrix dp_align per round."""
    b32 = np.asarray(b, dtype=np.float32)
    nb = b32.size - 1
    used = set(seed)
    count = 0
    thr = f32(thr)
    curr_best, exp_thresh = thr, thr
    out = []
    while curr_best >= exp_thresh and count < max_alns:
        S, prev = oracle.fill(b, x, bp, xp, 0, 0, 6, sorted(used) if used else None, 0)
        st = oracle.start(S, 0)
        oj, oq, os_ = oracle.traceback(S, prev, *st)
        first_lab, last_lab = int(oj[-1]), int(oj[0])
        by_lab = thr * (f32(last_lab - first_lab + 1) / f32(nb))
        by_len = thr * ((b32[last_lab] - b32[first_lab]) / b32[nb])
        exp_thresh = max(by_lab, by_len)
        curr_best = S[st]
        out.append((curr_best, st, oj, oq, os_))
        if curr_best > exp_thresh:
            count += 1
            used |= set(int(v) for v in oj)
        else:
            break
    return out


def make_pairs(rng, n_pairs, nb_range, nx_range):
    g = synth.genome_positions(rng, 30_000)
    contigs, segs = {}, {}
    for k in range(n_pairs):
        nb = int(rng.integers(*nb_range))
        nx = int(rng.integers(*nx_range))
        s0 = int(rng.integers(10, 29_000 - nb))
        contigs[k + 1] = synth.noisy_contig(rng, g, s0, nb)
        # the segment lies inside the contig's window (a real alignment exists), sometimes twice over (repeat -> later rounds)
        segs[k + 1] = synth.window_map(g, s0 + int(rng.integers(0, max(nb - nx, 1))), nx)
        if k % 5 == 0:
            contigs[k + 1] = np.round(contigs[k + 1])  # integer positions: exact ties
    return contigs, segs


def upload(ctx, maps, probs, side):
    ids = sorted(maps)
    off = np.zeros(len(ids) + 1, dtype=np.int64)
    for k, i in enumerate(ids):
        off[k + 1] = off[k] + len(maps[i])
    ctx.set_maps(side, off, np.concatenate([maps[i] for i in ids]).astype(np.float32),
                 np.concatenate([probs[i] for i in ids]).astype(np.float32))
    return ids


def check(ctx, oracle, contigs, segs, cids, sids, cp, sp, pr, thr, max_alns, seeds, env_note=""):
    words = (max(len(contigs[c]) for c in cids) + 31) // 32
    used_bits = None
    if seeds is not None:
        used_bits = np.zeros((pr.size, words), dtype=np.uint32)
        for k, u in enumerate(seeds):
            for j in u:
                used_bits[k, j >> 5] |= np.uint32(1 << (j & 31))
        pr = pr.copy()
        pr["used_slot"] = np.arange(pr.size)
    mode = ctx.mode(capi.SA_INIT_SEMIGLOBAL, capi.SA_START_SEMIGLOBAL, 0, 6, 0)
    nr, rounds, pj, pq, ps = ctx.align_rounds(mode, pr, thr, max_alns, None if used_bits is None else used_bits.ravel(), words)
    n_multi = 0
    for k in range(pr.size):
        c, s = int(pr["contig"][k]), int(pr["seg"][k])
        exp = reference_rounds(oracle, contigs[cids[c]], segs[sids[s]], cp[cids[c]], sp[sids[s]], thr[k], max_alns,
                               seeds[k] if seeds is not None else [])
        assert nr[k] == len(exp), f"pair {k}{env_note}: {nr[k]} rounds on the device, {len(exp)} in the reference loop"
        n_multi += len(exp) > 1
        for r, (best, st, oj, oq, os_) in enumerate(exp):
            rec = rounds[k, r]
            assert (int(rec["j"]), int(rec["q"])) == st, f"pair {k} round {r}: start {(rec['j'], rec['q'])} vs {st}"
            assert bits(rec["score"]) == bits(best)
            L, o = int(rec["path_len"]), int(rec["path_off"])
            assert L == oj.size
            assert np.array_equal(pj[o:o + L], oj) and np.array_equal(pq[o:o + L], oq)
            assert np.array_equal(bits(ps[o:o + L]), bits(os_))
    return n_multi


@pytest.mark.parametrize("gen", [1, 2])
def test_rounds_match_full_matrix_reference_loop(ctx, oracle, gen):
    rng = np.random.default_rng(500 + gen)
    contigs, segs = make_pairs(rng, 36, (40, 420), (4, 45))
    cp, sp = oracle.probs(contigs, gen), oracle.probs(segs, gen)
    cids = upload(ctx, contigs, cp, capi.SA_SIDE_CONTIG)
    sids = upload(ctx, segs, sp, capi.SA_SIDE_SEG)
    # every contig against its own segment and two others
    cc = np.repeat(np.arange(len(cids)), 3)
    ss = (cc + np.tile([0, 7, 19], len(cids))) % len(sids)
    pr = ctx.problems(cc, ss)
    mode = ctx.mode(capi.SA_INIT_SEMIGLOBAL, capi.SA_START_SEMIGLOBAL, 0, 6, 0)
    first = ctx.score_batch(mode, pr)["score"]
    thr = (first * rng.uniform(0.05, 1.6, size=pr.size)).astype(np.float32)
    thr[3] = np.nan        # NaN threshold: the loop never starts
    thr[5] = first[5]      # best == thr on round 1: exp_thresh <= thr, so round 1 is accepted unless the path spans the contig
    n_multi = check(ctx, oracle, contigs, segs, cids, sids, cp, sp, pr, thr, 6, None)
    assert n_multi > 20, "too few pairs ran more than one round: the windowed re-sweep is not exercised"
    # tip-phase form: seed bitmaps (labels already taken), gen-2 standard cap of 12 rounds
    seeds = [sorted(set(rng.integers(0, len(contigs[cids[int(c)]]) - 1, size=int(rng.integers(0, 12))).tolist())) for c in pr["contig"]]
    check(ctx, oracle, contigs, segs, cids, sids, cp, sp, pr, thr, 12, seeds)
    # whole runs of labels taken (what the tip pass inherits from the standard pass): 32-row strips that are entirely used are
    # not swept by the kernel but filled in with their known values -- the first strip (row 0), the ragged last strip, runs
    # that end one label short of a strip, and everything used
    seeds = []
    for k, c in enumerate(pr["contig"]):
        nb = len(contigs[cids[int(c)]]) - 1
        kind = k % 6
        if kind == 0:
            lo, hi = 0, min(nb, 32 * int(rng.integers(1, 4)))
        elif kind == 1:
            lo, hi = 32 * int(rng.integers(0, max(nb // 64, 1))), nb
        elif kind == 2:
            lo = 32 * int(rng.integers(0, max(nb // 32, 1)))
            hi = min(nb, lo + 32 * int(rng.integers(1, 5)) - 1)
        elif kind == 3:
            lo, hi = 0, nb
        elif kind == 4:
            lo = int(rng.integers(0, nb))
            hi = min(nb, lo + int(rng.integers(20, 200)))
        else:
            lo, hi = 32, min(nb, 64)
        s = set(range(lo, hi))
        if kind == 5:
            s |= set(range(min(nb, 96), min(nb, 160)))
        seeds.append(sorted(s))
    check(ctx, oracle, contigs, segs, cids, sids, cp, sp, pr, thr, 12, seeds)


def test_rounds_scratch_and_arena_limits(ctx, oracle, monkeypatch):
    """Pairs whose matrix exceeds the per-worker scratch come back with n_rounds = -1 (the caller falls back to
    sa_align_batch_packed rounds); a small path arena splits the batch into chunks without changing any result."""
    rng = np.random.default_rng(77)
    contigs, segs = make_pairs(rng, 24, (60, 300), (6, 30))
    cp, sp = oracle.probs(contigs, 2), oracle.probs(segs, 2)
    cids = upload(ctx, contigs, cp, capi.SA_SIDE_CONTIG)
    sids = upload(ctx, segs, sp, capi.SA_SIDE_SEG)
    pr = ctx.problems(np.arange(24), np.arange(24))
    mode = ctx.mode(capi.SA_INIT_SEMIGLOBAL, capi.SA_START_SEMIGLOBAL, 0, 6, 0)
    first = ctx.score_batch(mode, pr)["score"]
    thr = (first * 0.2).astype(np.float32)
    base = ctx.align_rounds(mode, pr, thr, 6)
    monkeypatch.setenv("SA_ROUNDS_ARENA_BYTES", "4000")
    small = ctx.align_rounds(mode, pr, thr, 6)
    assert np.array_equal(base[0], small[0])
    for k in range(24):
        for r in range(base[0][k]):
            a, b = base[1][k, r], small[1][k, r]
            assert (a["j"], a["q"], a["path_len"]) == (b["j"], b["q"], b["path_len"]) and bits(a["score"]) == bits(b["score"])
            assert np.array_equal(base[2][a["path_off"]:a["path_off"] + a["path_len"]], small[2][b["path_off"]:b["path_off"] + b["path_len"]])
    monkeypatch.delenv("SA_ROUNDS_ARENA_BYTES")
    cells = np.array([(len(contigs[cids[k]]) - 1) * (len(segs[sids[k]]) - 1) for k in range(24)])
    cut = float(np.median(cells)) + 0.25
    workers = 8 * ((24 + 7) // 8)
    monkeypatch.setenv("SA_ROUNDS_SCRATCH_BYTES", str((cut + 65536.0) * 2.2 * workers))
    nr = ctx.align_rounds(mode, pr, thr, 6)[0]
    assert np.array_equal(nr == -1, cells > cut), "the scratch cap did not reject exactly the oversized pairs"
    assert np.array_equal(nr[cells <= cut], base[0][cells <= cut])


@pytest.mark.parametrize("tip", [0, 1])
def test_filtered_rounds_return_exactly_the_writable_tracebacks(ctx, oracle, tip):
    """sa_align_rounds_filtered: same round records as sa_align_rounds (incl. j_first = first contig label of the path); a
    traceback comes back iff the round was accepted and passes the end / length / mean-score filters of run_SA_aln
    (/root/reference/SegAligner/main.cpp:303-325), and then it is the same traceback."""
    rng = np.random.default_rng(900 + tip)
    contigs, segs = make_pairs(rng, 40, (40, 420), (4, 45))
    cp, sp = oracle.probs(contigs, 2), oracle.probs(segs, 2)
    cids = upload(ctx, contigs, cp, capi.SA_SIDE_CONTIG)
    sids = upload(ctx, segs, sp, capi.SA_SIDE_SEG)
    cc = np.repeat(np.arange(len(cids)), 3)
    ss = (cc + np.tile([0, 5, 11], len(cids))) % len(sids)
    pr = ctx.problems(cc, ss)
    mode = ctx.mode(capi.SA_INIT_SEMIGLOBAL, capi.SA_START_SEMIGLOBAL, 0, 6, 0)
    first = ctx.score_batch(mode, pr)["score"]
    thr = (first * rng.uniform(0.05, 1.2, size=pr.size)).astype(np.float32)
    min_len, mean_min, mean_short = 4, 2500.0, 3500.0  # looser than run_SA_aln's constants so that both outcomes are frequent here
    flt = capi.RoundsFilter(tip, min_len, mean_min, mean_short)
    nr0, r0, pj0, pq0, ps0 = ctx.align_rounds(mode, pr, thr, 6)
    nr1, r1, pj1, pq1, ps1 = ctx.align_rounds(mode, pr, thr, 6, flt=flt)
    assert np.array_equal(nr0, nr1)
    kept = dropped = 0
    for k in range(pr.size):
        nb = len(contigs[cids[int(pr["contig"][k])]]) - 1
        cpos = np.asarray(contigs[cids[int(pr["contig"][k])]], dtype=np.float32)
        for r in range(nr0[k]):
            a, b = r0[k, r], r1[k, r]
            assert (a["j"], a["q"], a["path_len"], a["j_first"]) == (b["j"], b["q"], b["path_len"], b["j_first"])
            assert bits(a["score"]) == bits(b["score"])
            L, o = int(a["path_len"]), int(a["path_off"])
            assert int(a["j_first"]) == int(pj0[o + L - 1]) and int(a["j"]) == int(pj0[o])
            first_lab, last_lab = int(pj0[o + L - 1]), int(pj0[o])
            by_lab = np.float32(thr[k]) * (np.float32(last_lab - first_lab + 1) / np.float32(nb))
            by_len = np.float32(thr[k]) * ((cpos[last_lab] - cpos[first_lab]) / cpos[nb])
            accepted = a["score"] > max(by_lab, by_len)
            mean = float(a["score"]) / max(L - 1, 1)
            bar = mean_short if L < 5 else mean_min
            want = bool(accepted and L >= min_len and mean >= 0.999 * bar)
            if tip and want and (nb - 1 - last_lab > 3 and first_lab > 2):
                want = False
            sure = abs(mean / bar - 0.999) > 1e-3  # the mean test carries 0.1 % slack: only check clear cases
            if sure or not accepted:
                assert (int(b["path_off"]) >= 0) == want, (k, r, accepted, L, mean, first_lab, last_lab, nb)
            if int(b["path_off"]) >= 0:
                kept += 1
                o1 = int(b["path_off"])
                assert np.array_equal(pj1[o1:o1 + L], pj0[o:o + L]) and np.array_equal(pq1[o1:o1 + L], pq0[o:o + L])
                assert np.array_equal(bits(ps1[o1:o1 + L]), bits(ps0[o:o + L]))
            else:
                dropped += 1
    assert kept > 5 and dropped > 5, (kept, dropped)


def test_stop_below_truncates_the_loop_after_the_first_low_round(ctx, oracle):
    """sa_rounds_filter.stop_below: the pair's loop ends after the first round that scores below it (that round is still
    reported); everything up to there is the unfiltered loop's, and round scores never increase (what makes the cut exact)."""
    rng = np.random.default_rng(4242)
    contigs, segs = make_pairs(rng, 40, (40, 420), (4, 45))
    cp, sp = oracle.probs(contigs, 2), oracle.probs(segs, 2)
    cids = upload(ctx, contigs, cp, capi.SA_SIDE_CONTIG)
    sids = upload(ctx, segs, sp, capi.SA_SIDE_SEG)
    cc = np.repeat(np.arange(len(cids)), 3)
    ss = (cc + np.tile([0, 5, 11], len(cids))) % len(sids)
    pr = ctx.problems(cc, ss)
    mode = ctx.mode(capi.SA_INIT_SEMIGLOBAL, capi.SA_START_SEMIGLOBAL, 0, 6, 0)
    first = ctx.score_batch(mode, pr)["score"]
    thr = (first * 0.05).astype(np.float32)
    nr0, r0, pj0, pq0, ps0 = ctx.align_rounds(mode, pr, thr, 6)
    scores = np.array([r0[k, r]["score"] for k in range(pr.size) for r in range(nr0[k])])
    stop = float(np.median(scores))
    flt = capi.RoundsFilter(0, 2, 0.0, 0.0, stop, 0)
    nr1, r1, pj1, pq1, ps1 = ctx.align_rounds(mode, pr, thr, 6, flt=flt)
    cut = 0
    for k in range(pr.size):
        sc = [float(r0[k, r]["score"]) for r in range(nr0[k])]
        assert all(a >= b for a, b in zip(sc, sc[1:])), f"pair {k}: round scores increase {sc}"
        below = [r for r, v in enumerate(sc) if v < stop]
        want = min(nr0[k], below[0] + 1) if below else nr0[k]
        assert nr1[k] == want, (k, sc, stop, nr1[k])
        cut += want < nr0[k]
        for r in range(want):
            a, b = r0[k, r], r1[k, r]
            assert (a["j"], a["q"], a["path_len"], a["j_first"]) == (b["j"], b["q"], b["path_len"], b["j_first"])
            assert bits(a["score"]) == bits(b["score"])
    assert cut > 10, cut


def test_tip_windows_only_drop_pairs_that_cannot_write_a_tip_alignment(ctx, oracle):
    """sa_rounds_filter.tip_windows: a pair reported with n_rounds = -2 has, in its complete loop, no round that touches a
    contig end, is long enough and passes the mean-score filter (so nothing it could have written); every other pair's rounds
    are the truncated loop's."""
    rng = np.random.default_rng(8080)
    g = synth.genome_positions(rng, 30_000)
    contigs, segs = {}, {}
    for k in range(48):
        nb = int(rng.integers(220, 460))
        nx = int(rng.integers(4, 15))
        s0 = int(rng.integers(10, 29_000 - nb))
        contigs[k + 1] = synth.noisy_contig(rng, g, s0, nb)
        at = [0, nb - nx, int(rng.integers(0, nb - nx)), int(rng.integers(40, nb - nx - 40))][k % 4]  # top end, bottom end, anywhere, inside
        segs[k + 1] = synth.window_map(g, s0 + at, nx)
    cp, sp = oracle.probs(contigs, 2), oracle.probs(segs, 2)
    cids = upload(ctx, contigs, cp, capi.SA_SIDE_CONTIG)
    sids = upload(ctx, segs, sp, capi.SA_SIDE_SEG)
    cc = np.repeat(np.arange(len(cids)), 3)
    ss = (cc + np.tile([0, 5, 11], len(cids))) % len(sids)
    pr = ctx.problems(cc, ss)
    mode = ctx.mode(capi.SA_INIT_SEMIGLOBAL, capi.SA_START_SEMIGLOBAL, 0, 6, 0)
    first = ctx.score_batch(mode, pr)["score"]
    thr = (first * 0.05).astype(np.float32)
    # seed bitmaps as the tip pass has them: a run of labels taken somewhere inside the contig
    words = (max(len(contigs[c]) for c in cids) + 31) // 32
    used_bits = np.zeros((pr.size, words), dtype=np.uint32)
    for k in range(pr.size):
        nb = len(contigs[cids[int(pr["contig"][k])]]) - 1
        lo = int(rng.integers(0, nb - 30))
        for j in range(lo, lo + int(rng.integers(0, 90))):
            if j < nb and rng.random() < 0.8:
                used_bits[k, j >> 5] |= np.uint32(1 << (j & 31))
    pr = pr.copy()
    pr["used_slot"] = np.arange(pr.size)
    nr0, r0, pj0, pq0, ps0 = ctx.align_rounds(mode, pr, thr, 6, used_bits.ravel(), words)
    min_len, mean_min, mean_short = 5, 7000.0, 8000.0
    stop = 0.99 * (min_len - 1) * mean_min
    flt = capi.RoundsFilter(1, min_len, mean_min, mean_short, stop, 1)
    nr1, r1, pj1, pq1, ps1 = ctx.align_rounds(mode, pr, thr, 6, used_bits.ravel(), words, flt=flt)
    dead = alive = could_write = 0
    for k in range(pr.size):
        nb = len(contigs[cids[int(pr["contig"][k])]]) - 1
        touching = [(nb - 1 - int(r0[k, r]["j"]) <= 3 or int(r0[k, r]["j_first"]) <= 2) and int(r0[k, r]["path_len"]) >= min_len
                    and float(r0[k, r]["score"]) / (int(r0[k, r]["path_len"]) - 1) >= 0.998 * mean_min for r in range(nr0[k])]
        if nr1[k] == -2:
            dead += 1
            assert not any(touching), (k, [(float(r0[k, r]["score"]), int(r0[k, r]["j_first"]), int(r0[k, r]["j"])) for r in range(nr0[k])], nb)
            continue
        alive += 1
        could_write += any(touching)
        sc = [float(r0[k, r]["score"]) for r in range(nr0[k])]
        below = [r for r, v in enumerate(sc) if v < stop]
        want = min(nr0[k], below[0] + 1) if below else nr0[k]
        assert nr1[k] == want, (k, sc, nr1[k])
        for r in range(want):
            a, b = r0[k, r], r1[k, r]
            assert (a["j"], a["q"], a["path_len"], a["j_first"]) == (b["j"], b["q"], b["path_len"], b["j_first"])
            assert bits(a["score"]) == bits(b["score"])
            if int(b["path_off"]) >= 0:
                L, o0, o1 = int(a["path_len"]), int(a["path_off"]), int(b["path_off"])
                assert np.array_equal(pj1[o1:o1 + L], pj0[o0:o0 + L]) and np.array_equal(pq1[o1:o1 + L], pq0[o0:o0 + L])
                assert np.array_equal(bits(ps1[o1:o1 + L]), bits(ps0[o0:o0 + L]))
    assert dead > 10 and alive > 10 and could_write > 5, (dead, alive, could_write)
