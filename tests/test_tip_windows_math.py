"""CPU: the mathematics of the tip pass's end-window certificate (sa_rounds_filter.tip_windows) against the
oracle.  The certificate is restated here in numpy exactly as csrc/sa_capi.cu / sa_kernels.cu compute it -- window spans from
the skip-penalty budget, the top window with the free starts below row 2 closed, the bottom window swept cold, a per-step gain
G instead of score_f's 10000 (/root/reference/SegAligner/SegAligner.cpp:83-92) -- and checked on seeded pairs: whenever it
declares a pair dead, the pair's COMPLETE alignment loop (run_SA_aln, /root/reference/SegAligner/main.cpp:262-325, run with the
oracle's full matrices) has no round that touches a contig end, is long enough and passes the mean-score filter.  The GPU test
tests/test_gpu_rounds.py::test_tip_windows_only_drop_pairs_that_cannot_write_a_tip_alignment checks the kernel the same way."""
import math

import numpy as np
import pytest

import helpers
from ampliconreconstructorom_b200 import synth

f32 = np.float32
LB = 6


@pytest.fixture(scope="module")
def oracle():
    return helpers.Oracle()


def window_values(b, x, xprob, used, gain, row_lo, row_hi, closed_below_row2):
    """dp_align restricted to rows [row_lo, row_hi) with nothing above them, `gain` in place of 10000; returns S (rows, nx)."""
    b = np.asarray(b, dtype=f32)
    x = np.asarray(x, dtype=f32)
    xprob = np.asarray(xprob, dtype=f32)
    nx = x.size - 1
    S = np.full((row_hi - row_lo, nx), -np.inf, dtype=f32)
    for j in range(row_lo, row_hi):
        for q in range(nx):
            init = f32(0.0) if (j == 0 or (q < 2 and (not closed_below_row2 or j <= 2))) else f32(-np.inf)
            best = init
            if j not in used:
                for i in range(max(j - LB, row_lo, 0), j):
                    for p in range(max(q - LB, 0), q):
                        sc = S[i - row_lo, p]
                        if sc == -np.inf:
                            continue
                        exp_x = f32(0.0)
                        for k in range(p + 1, q):
                            exp_x = f32(exp_x + xprob[k])
                        discrep = f32(abs(f32(f32(b[j] - b[i]) - f32(x[q] - x[p]))))
                        delta = f32(math.pow(float(discrep), 1.2))  # the certificate has margins of tens: no need for glibc's bits
                        pen = f32(f32(f32(5000.0) * f32(j - i - 1) + f32(5000.0) * exp_x) + delta)
                        v = f32(sc + f32(f32(gain) - pen))
                        if v > best:
                            best = v
            S[j - row_lo, q] = best
    return S


def certificate_says_dead(b, x, xprob, used, min_len, mean_min, mean_short):
    """The pre-check as the library sets it up (sa_capi.cu: r_win_gain / r_win_min / r_win_slope; sa_kernels.cu: win_span,
    the two passes, end_window_best).  None when the two windows overlap (the pair is aligned as usual)."""
    nb, nx = len(b) - 1, len(x) - 1
    m = f32(0.999) * f32(min(mean_min, mean_short))
    c = f32(10000.0) - m
    gain = min(f32(10000.0), f32(1.1) * f32(min_len - 1) * c)
    w_min = (gain - c) * f32(min_len - 1) - f32(10.0)
    slope = (f32(1.0) + c / f32(5000.0)) * f32(1.001)
    span = min(int(math.ceil(float(f32(nx - 1) * slope))) + 1, LB * (nx - 1))
    s_top = (3 + span + 31) >> 5
    bot_row = nb - 4 - span
    if not (bot_row > 0 and (bot_row >> 5) >= s_top):
        return None
    top = window_values(b, x, xprob, used, gain, 0, min(nb, 32 * s_top), True)
    lo = 32 * (bot_row >> 5)
    bot = window_values(b, x, xprob, used, gain, lo, nb, False)
    best = -np.inf
    for j in range(min(nb, 3 + span)):
        best = max(best, float(top[j, nx - 2]), float(top[j, nx - 1]))
    for j in range(max(0, nb - 4), nb):
        best = max(best, float(bot[j - lo, nx - 2]), float(bot[j - lo, nx - 1]))
    best = max(best, float(bot[nb - 1 - lo].max()))
    return best < float(w_min)


def full_loop(oracle, b, x, bp, xp, thr, max_alns, seed):
    """run_SA_aln's loop with full matrices: [(score, path rows end -> start)] of every round it runs."""
    b32 = np.asarray(b, dtype=f32)
    nb = b32.size - 1
    used = set(seed)
    count = 0
    thr = f32(thr)
    curr_best, exp_thresh = thr, thr
    out = []
    while curr_best >= exp_thresh and count < max_alns:
        S, prev = oracle.fill(b, x, bp, xp, 0, 0, 6, sorted(used) if used else None, 0)
        st = oracle.start(S, 0)
        oj, oq, _ = oracle.traceback(S, prev, *st)
        first_lab, last_lab = int(oj[-1]), int(oj[0])
        exp_thresh = max(thr * (f32(last_lab - first_lab + 1) / f32(nb)), thr * ((b32[last_lab] - b32[first_lab]) / b32[nb]))
        curr_best = S[st]
        out.append((float(curr_best), oj))
        if curr_best > exp_thresh:
            count += 1
            used |= set(int(v) for v in oj)
        else:
            break
    return out


@pytest.mark.parametrize("gen", [1, 2])
def test_dead_pairs_have_no_writable_round(oracle, gen):
    rng = np.random.default_rng(31337 + gen)
    min_len = 3 if gen == 1 else 5  # main.cpp:218-222
    mean_min, mean_short = 8750.0, 9400.0
    g = synth.genome_positions(rng, 20_000)
    dead = alive = alive_writable = unchecked = 0
    for k in range(36):
        nb = int(rng.integers(150, 260))
        nx = int(rng.integers(5, 11))
        s0 = int(rng.integers(10, 19_000 - nb))
        b = synth.noisy_contig(rng, g, s0, nb) if k % 6 else synth.window_map(g, s0, nb)
        kind = k % 4
        if kind == 0:    # the segment hangs over the top end of the contig: a true tip alignment
            x = synth.window_map(g, s0 - 3, nx + 3)
        elif kind == 1:  # ... over the bottom end
            x = synth.window_map(g, s0 + nb - nx, nx + 3)
        elif kind == 2:  # inside the contig: a true alignment that touches no end
            x = synth.window_map(g, s0 + int(rng.integers(60, nb - 60 - nx)), nx)
        else:            # from elsewhere: chance alignments only
            x = synth.window_map(g, int(rng.integers(10, 19_000)), nx)
        maps = {1: b}
        bp = oracle.probs(maps, gen)[1]
        xp = oracle.probs({1: x}, gen)[1]
        nb_l = len(b) - 1
        used = set()
        lo = int(rng.integers(20, nb_l - 40))
        for j in range(lo, lo + int(rng.integers(0, 60))):
            if j < nb_l and rng.random() < 0.8:
                used.add(j)
        verdict = certificate_says_dead(b, x, xp, used, min_len, mean_min, mean_short)
        if verdict is None:
            unchecked += 1
            continue
        rounds = full_loop(oracle, b, x, bp, xp, 1000.0, 6, used)
        writable = []
        for score, oj in rounds:
            n = oj.size
            touches = (nb_l - 1 - int(oj[0]) <= 3) or int(oj[-1]) <= 2
            bar = mean_short if n < 5 else mean_min
            writable.append(touches and n >= min_len and score / (n - 1) >= 0.999 * bar)
        if verdict:
            dead += 1
            assert not any(writable), (k, kind, [(s, int(oj[-1]), int(oj[0]), oj.size) for s, oj in rounds], nb_l)
        else:
            alive += 1
            alive_writable += any(writable)
    print(f"gen {gen}: {dead} dead, {alive} alive ({alive_writable} of them with a writable round), {unchecked} unchecked")
    assert dead >= 8 and alive_writable >= 4 and unchecked == 0, (dead, alive, alive_writable, unchecked)
