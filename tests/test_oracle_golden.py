"""CPU: pins the oracle (oracle/sa_oracle.c) and the host-side product code against
  (1) the committed golden vectors produced by the unmodified reference (tests/golden/ka_vectors.json, made by
      tests/golden/make_golden.py; they include the KA-5..KA-7 known answers of SURVEY.md 8(c)), and
  (2) the reference itself through oracle/_ref/libsa_ref.so when that library is present (build container)."""
import hashlib
import json

import numpy as np
import pytest

import helpers
from ampliconreconstructorom_b200 import capi

G = json.loads((helpers.GOLDEN / "ka_vectors.json").read_text())


def fh(lst):
    return np.array([float.fromhex(v) for v in lst], dtype=np.float32)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


X, B = fh(G["x"]), fh(G["b"])


def test_survey_known_answers_are_in_the_fixture():
    # SURVEY.md 8(c) KA-5/6/7 literal values
    assert G["probs_x_gen1"][2] == float.fromhex("0x1.61e1cep-5").hex()
    assert G["probs_x_gen1"][-1] == float.fromhex("0x1.57c526p-6").hex()
    assert G["score_f"][2]["plain"] == float.fromhex("-0x1.7dc9cep+15").hex()
    assert G["score_f"][2]["swapped"] == float.fromhex("-0x1.a329d2p+15").hex()
    d0 = G["dp_align"][0]
    assert d0["start"] == [8, 8] and d0["path_s"][0] == float.fromhex("0x1.0e7408p+16").hex()
    assert d0["path_j"] == [8, 7, 6, 5, 4, 3, 2, 1, 0]
    d1 = G["dp_align"][1]
    assert d1["path_j"] == [8, 7, 6, 5, 4, 2, 1, 0] and d1["path_s"][0] == float.fromhex("0x1.a65b44p+15").hex()


def test_oracle_reverse_and_probs(oracle):
    assert np.array_equal(bits(oracle.reverse(X)), bits(fh(G["reverse_x"])))
    for gen in (1, 2):
        assert np.array_equal(bits(oracle.probs({1: X}, gen)[1]), bits(fh(G[f"probs_x_gen{gen}"])))
        assert np.array_equal(bits(oracle.probs({1: B}, gen)[1]), bits(fh(G[f"probs_b_gen{gen}"])))


def test_host_code_reverse_and_probs():
    """the product's host C++ (csrc/host/cmap.cpp) through the C ABI"""
    ids, off, pos = capi.flatten_maps({1: X})
    ri, ro, rp = capi.host_make_reverse(ids, off, pos, 10)
    full = capi.unflatten_maps(ri, ro, rp)
    assert sorted(full) == [-1, 1]
    assert np.array_equal(bits(full[-1]), bits(fh(G["reverse_x"]))) and np.array_equal(bits(full[1]), bits(X))
    assert capi.host_make_reverse(ids, off, pos, 14)[0].size == 0  # size() <= min_map_len is dropped
    for gen in (1, 2):
        for name, m in (("x", X), ("b", B)):
            _, o, p = capi.flatten_maps({1: m})
            assert np.array_equal(bits(capi.host_non_collapse_probs(o, p, gen)), bits(fh(G[f"probs_{name}_gen{gen}"])))


def test_oracle_score_f(oracle):
    px, pb = fh(G["probs_x_gen1"]), fh(G["probs_b_gen1"])
    for e in G["score_f"]:
        i, j, p, q = e["ijpq"]
        assert bits(oracle.score_f(B, X, i, j, p, q, px)) == bits(float.fromhex(e["plain"]))
        assert bits(oracle.score_f(X, B, p, q, i, j, pb)) == bits(float.fromhex(e["swapped"]))


def test_oracle_dp_align_known_answers(oracle):
    px, pb = fh(G["probs_x_gen1"]), fh(G["probs_b_gen1"])
    for d in G["dp_align"]:
        S, prev = oracle.fill(B, X, pb, px, d["init"], d["fitting"], 6, d["used"], d["swap"])
        assert np.array_equal(bits(S).ravel(), bits(fh(d["S"])))
        assert prev.ravel().tolist() == d["prev"]
        st = oracle.start(S, d["start_mode"])
        assert list(st) == d["start"]
        pj, pq, ps = oracle.traceback(S, prev, *st)
        assert pj.tolist() == d["path_j"] and pq.tolist() == d["path_q"] and np.array_equal(bits(ps), bits(fh(d["path_s"])))


def test_oracle_random_cases_match_reference_digests(oracle):
    rng = np.random.default_rng(2024)
    for r in G["random"]:
        nb, nx = int(rng.integers(20, 140)), int(rng.integers(8, 60))
        assert (nb, nx) == (r["nb"], r["nx"])
        bb, xx = helpers.related_maps(rng, nb, nx, integer=r["integer"])
        pbb, pxx = oracle.probs({1: bb}, r["gen"])[1], oracle.probs({1: xx}, r["gen"])[1]
        init, fit, swap, start = r["mode"]
        S, prev = oracle.fill(bb, xx, pbb, pxx, init, fit, 6, None, swap)
        assert hashlib.sha256(S.tobytes()).hexdigest() == r["S_sha256"]
        assert hashlib.sha256(prev.tobytes()).hexdigest() == r["prev_sha256"]
        assert list(oracle.start(S, start)) == r["start"]
        if r["multi_backtrack_swap"]:
            S2, prev2 = oracle.fill(bb, xx, pbb, pxx, 0, 0, 6, None, 1)
            mb = oracle.multi_backtrack(S2, prev2, 4)
            assert np.array_equal(bits(mb), bits(fh(r["multi_backtrack_swap"])))


def test_oracle_against_live_reference(oracle, ref):
    """Differential test against the reference's own functions (only where oracle/_ref is built)."""
    rng = np.random.default_rng(77)
    for t in range(30):
        nb, nx = int(rng.integers(2, 90)), int(rng.integers(2, 50))
        b, x = helpers.related_maps(rng, nb, nx, integer=(t % 3 == 0))
        gen = 1 + t % 2
        maps = {1: b, 2: x}
        po, pr = oracle.probs(maps, gen), ref.probs(maps, gen)
        assert all(np.array_equal(bits(po[k]), bits(pr[k])) for k in po)
        for (init, fit, swap, start) in [(0, 0, 0, 0), (1, 0, 1, 1), (2, 1, 0, 2), (0, 0, 1, 1)]:
            used = sorted(set(rng.integers(0, nb, size=3).tolist())) if t % 4 == 1 and not fit else None
            S, prev = oracle.fill(b, x, po[1], po[2], init, fit, 6, used, swap)
            S2, prev2, st2, path2 = ref.dp(b, x, po[1], po[2], init, fit, 6, used, swap, start)
            assert np.array_equal(bits(S), bits(S2)) and np.array_equal(prev, prev2)
            st = oracle.start(S, start)
            assert st == st2
            if st[0] >= 0:
                pj, pq, ps = oracle.traceback(S, prev, *st)
                assert np.array_equal(pj, path2[0]) and np.array_equal(pq, path2[1]) and np.array_equal(bits(ps), bits(path2[2]))


def test_oracle_unbanded_lookback_against_reference(oracle, ref):
    """-nl: lookback = 9999999 (main.cpp:475-478)"""
    rng = np.random.default_rng(5)
    for t in range(4):
        b, x = helpers.related_maps(rng, int(rng.integers(8, 40)), int(rng.integers(6, 25)), integer=(t == 0))
        po = oracle.probs({1: b, 2: x}, 2)
        S, prev = oracle.fill(b, x, po[1], po[2], 0, 0, 9999999, None, 0)
        S2, prev2, _, _ = ref.dp(b, x, po[1], po[2], 0, 0, 9999999, None, 0, 0)
        assert np.array_equal(bits(S), bits(S2)) and np.array_equal(prev, prev2)
