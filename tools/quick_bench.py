"""Developer probe (not the bench contract): scoring-phase GCUPS of the fill kernel on a C2-shaped sample."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from ampliconreconstructorom_b200 import capi, synth
import helpers

n_segs = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_contigs = int(sys.argv[2]) if len(sys.argv) > 2 else 100
swap = int(sys.argv[3]) if len(sys.argv) > 3 else 0
o = helpers.Oracle()
uniform = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if uniform:  # equal-size problems: steady-state kernel efficiency without any tail
    segs, contigs = synth.make_workload(2000, n_segs, n_contigs, seg_clip=(64, 64), contig_clip=(1024, 1024))
else:
    segs, contigs = synth.make_workload(2000, n_segs, n_contigs)
full = {}
for i, m in segs.items():
    full[i] = m.astype(np.float32); full[-i] = o.reverse(m.astype(np.float32))
contigs = {i: m.astype(np.float32) for i, m in contigs.items()}
sp = o.probs(full, 2); cp = o.probs(contigs, 2)
ctx = capi.Context(0)
def up(maps, probs, side):
    ids = sorted(maps); off = np.zeros(len(ids) + 1, dtype=np.int64)
    for k, i in enumerate(ids): off[k + 1] = off[k] + len(maps[i])
    ctx.set_maps(side, off, np.concatenate([maps[i] for i in ids]), np.concatenate([probs[i] for i in ids]))
    return ids
cids = up(contigs, cp, 0); sids = up(full, sp, 1)
cc, ss = np.meshgrid(np.arange(len(cids)), np.arange(len(sids)), indexing="ij")
pr = ctx.problems(cc.ravel(), ss.ravel())
nb = ctx.n_labels[0][pr["contig"]]; nx = ctx.n_labels[1][pr["seg"]]
cells = float((nb * nx).sum())
# pair evaluations (SURVEY 8d work model)
def tri(n): j = np.arange(n); return np.minimum(j, 6).sum()
Eb = np.array([tri(n) for n in ctx.n_labels[0]]); Ex = np.array([tri(n) for n in ctx.n_labels[1]])
evals = float((Eb[pr["contig"]] * Ex[pr["seg"]]).sum())
print(f"problems={pr.size} cells={cells:.3e} evals/cell={evals/cells:.2f}")
peak = ctx.fp64_peak(); print(f"fp64 peak (DFMA thread-instr/s) = {peak:.3e}")
mode = ctx.mode(0, 0, 0, 6, swap)
plan = ctx.plan_create(pr); print(plan.info())
for rep in range(3):
    t0 = time.time(); plan.score(mode); res = plan.results(); t1 = time.time()
    tm = ctx.timing()
    print(f"rep{rep}: fill {tm.fill_ms:.1f} ms  -> {cells/tm.fill_ms/1e6:.2f} GCUPS kernel, e2e {cells/(t1-t0)/1e9:.2f} GCUPS; "
          f"fp64 frac (18/eval) = {18*evals/(tm.fill_ms*1e-3)/peak:.3f}")
# spot check vs oracle
rng = np.random.default_rng(0)
for k in rng.choice(pr.size, size=5, replace=False):
    c, s = int(pr["contig"][k]), int(pr["seg"][k])
    So, _ = o.fill(contigs[cids[c]], full[sids[s]], cp[cids[c]], sp[sids[s]], 0, 0, 6, None, swap)
    st = o.start(So, 0)
    ok = (int(res["j"][k]), int(res["q"][k])) == st and np.float32(res["score"][k]).view(np.uint32) == So[st].view(np.uint32)
    print("check", k, c, s, "OK" if ok else f"MISMATCH {res[k]} vs {st} {So[st]}")
