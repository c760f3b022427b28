"""Regenerates tests/golden/*.json from the UNMODIFIED reference (run in the build container, where
/root/reference exists and oracle/_ref has been built by `make -C oracle`).

  ka_vectors.json   unit-level known answers: make_reverse_cmap, non_collapse_probs, score_f, dp_align
                    (inputs of SURVEY.md 8(c) KA-5..KA-7 plus seeded random cases), floats as C hex strings,
                    produced by calling the reference's own functions through oracle/_ref/libsa_ref.so.
  cli_digests.json  sha256 of every output file of the reference SegAligner binary on the CLI cases of
                    tests/cli_cases.py (inputs are regenerated from seeds at test time).
"""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import cli_cases as cc  # noqa: E402
import helpers  # noqa: E402


def hx(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float32).ravel()]


def main():
    ref = helpers.RefShim.try_load()
    assert ref is not None, "build oracle/_ref first (make -C oracle)"
    out = {}
    # KA-5..7 inputs
    x = np.array([614, 8539, 39211, 40100, 43539, 51249, 75750, 75764, 77134, 89880, 90107, 96798, 98001, 101273],
                 dtype=np.float32)
    b = np.array([1000, 8950.5, 39700.2, 40600, 44010.7, 51800, 76200.3, 77600.9, 80000, 90000], dtype=np.float32)
    out["x"], out["b"] = hx(x), hx(b)
    rev = ref.make_reverse({1: x}, 10)
    out["reverse_x"] = hx(rev[-1])
    for gen in (1, 2):
        out[f"probs_x_gen{gen}"] = hx(ref.probs({1: x}, gen)[1])
        out[f"probs_b_gen{gen}"] = hx(ref.probs({1: b}, gen)[1])
    px, pb = ref.probs({1: x}, 1)[1], ref.probs({1: b}, 1)[1]
    sf = []
    for (i, j, p, q) in [(0, 1, 0, 1), (0, 2, 0, 1), (1, 2, 0, 3), (2, 6, 2, 8), (0, 5, 1, 4), (3, 4, 3, 4)]:
        sf.append({"ijpq": [i, j, p, q], "plain": float(np.float32(ref.score_f(b, x, i, j, p, q, px))).hex(),
                   "swapped": float(np.float32(ref.score_f(x, b, p, q, i, j, pb))).hex()})
    out["score_f"] = sf
    dp = []
    for (init, fit, swap, start, used) in [(0, 0, 0, 0, None), (0, 0, 0, 0, [3]), (1, 0, 1, 1, None), (2, 1, 0, 2, None),
                                           (0, 0, 1, 1, None)]:
        S, prev, st, path = ref.dp(b, x, pb, px, init, fit, 6, used, swap, start)
        dp.append({"init": init, "fitting": fit, "swap": swap, "start_mode": start, "used": used, "start": list(st),
                   "S": hx(S), "prev": prev.ravel().tolist(), "path_j": path[0].tolist(), "path_q": path[1].tolist(),
                   "path_s": hx(path[2])})
    out["dp_align"] = dp
    # seeded random cases (bigger, multi-strip): only digests of S/prev to keep the fixture small
    import hashlib
    rnd = []
    rng = np.random.default_rng(2024)
    for t in range(12):
        nb, nx = int(rng.integers(20, 140)), int(rng.integers(8, 60))
        bb, xx = helpers.related_maps(rng, nb, nx, integer=(t % 3 == 0))
        gen = 1 + t % 2
        pbb, pxx = ref.probs({1: bb}, gen)[1], ref.probs({1: xx}, gen)[1]
        init, fit, swap, start = [(0, 0, 0, 0), (0, 0, 1, 1), (1, 0, 1, 1), (2, 1, 0, 2)][t % 4]
        S, prev, st, path = ref.dp(bb, xx, pbb, pxx, init, fit, 6, None, swap, start)
        n_mb = 4
        mb = ref.multi_backtrack(bb, xx, pbb, pxx, n_mb, 6, 1) if int((S > 0).sum()) > n_mb + 1 else np.zeros(0, np.float32)
        rnd.append({"seed_index": t, "nb": nb, "nx": nx, "integer": t % 3 == 0, "gen": gen, "mode": [init, fit, swap, start],
                    "S_sha256": hashlib.sha256(S.tobytes()).hexdigest(), "prev_sha256": hashlib.sha256(prev.tobytes()).hexdigest(),
                    "start": list(st), "path_len": int(path[0].size), "score": float(path[2][0]).hex() if path[2].size else None,
                    "multi_backtrack_swap": hx(mb)})
    out["random"] = rnd
    (HERE / "ka_vectors.json").write_text(json.dumps(out, indent=0))

    dig = {}
    for name, (kw, flags) in cc.CASES.items():
        segs, contigs = cc.make_case(**kw)
        with tempfile.TemporaryDirectory() as d:
            d = Path(d)
            s, c = cc.write_case(d / "in", segs, contigs)
            r = cc.run_bin(cc.REF_BIN, d / "ref", s, c, flags)
            assert r.returncode == 0
            dig[name] = cc.dir_digest(d / "ref")
    (HERE / "cli_digests.json").write_text(json.dumps(dig, indent=0, sort_keys=True))
    print("wrote", HERE / "ka_vectors.json", HERE / "cli_digests.json")


if __name__ == "__main__":
    main()
