"""Developer probe: coupled units vs one warp per problem on hand-picked shapes (prints which shapes disagree or time out)."""
import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from ampliconreconstructorom_b200 import capi
import helpers
orc = helpers.Oracle()
rng = np.random.default_rng(11)
ctx = capi.Context(0)


def run(shapes, tag):
    contigs, segs = {}, {}
    for k, (nb, nx) in enumerate(shapes):
        b, x = helpers.related_maps(rng, nb, nx)
        contigs[k + 1], segs[k + 1] = b, x
    cp, sp = orc.probs(contigs, 2), orc.probs(segs, 2)
    def up(maps, probs, side):
        ids = sorted(maps); off = np.zeros(len(ids) + 1, dtype=np.int64)
        for n, i in enumerate(ids): off[n + 1] = off[n] + len(maps[i])
        ctx.set_maps(side, off, np.concatenate([maps[i] for i in ids]).astype(np.float32), np.concatenate([probs[i] for i in ids]).astype(np.float32))
    up(contigs, cp, 0); up(segs, sp, 1)
    n = len(shapes)
    pr = ctx.problems(np.arange(n), np.arange(n))
    mode = ctx.mode(0, 0, 0, 6, 0)
    os.environ["SA_SHAPES"] = "coupled"; os.environ["SA_CPL_ALPHA"] = "0"
    ref = ctx.score_batch(mode, pr)
    os.environ["SA_CPL_ALPHA"] = "1e-9"
    t0 = time.time()
    try:
        got = ctx.score_batch(mode, pr)
        ok = got.tobytes() == ref.tobytes()
        bad = [k for k in range(n) if got[k].tobytes() != ref[k].tobytes()]
        print(f"{tag}: {n} problems, {time.time()-t0:.2f} s, equal={ok} bad={bad[:8]}", flush=True)
    except Exception as e:
        print(f"{tag}: {n} problems, {time.time()-t0:.2f} s, ERROR {e}", flush=True)


shapes = [(int(rng.integers(20, 300)), int(rng.integers(4, 120))) for _ in range(1600)]
only_multi = [sh for sh in shapes if sh[0] > 64 and sh[1] >= 8]
for cap in ("100", "148", "149", "200", "296", "297", "400", "444"):
    os.environ["SA_CPL_MAX_CTAS"] = cap
    print("== SA_CPL_MAX_CTAS =", cap)
    rng = np.random.default_rng(5)
    run(shapes, "1600 mixed problems")
    rng = np.random.default_rng(5)
    run(only_multi, "only multi-strip problems (no bulk kernel beside)")
