"""Developer probe (not the bench contract): scoring-phase GCUPS of the fill kernel on a C2-shaped sample."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import os
from ampliconreconstructorom_b200 import capi, synth
if os.environ.get("SA_LIB"):  # developer A/B builds from tools/build_variants.sh
    capi.LIB_PATH = Path(os.environ["SA_LIB"]).resolve()

n_segs = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_contigs = int(sys.argv[2]) if len(sys.argv) > 2 else 100
swap = int(sys.argv[3]) if len(sys.argv) > 3 else 0
uniform = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if uniform:  # equal-size problems: steady-state kernel efficiency without any tail
    segs, contigs = synth.make_workload(2000, n_segs, n_contigs, seg_clip=(64, 64), contig_clip=(1024, 1024))
else:
    segs, contigs = synth.make_workload(2000, n_segs, n_contigs)
sid, soff, spos = capi.flatten_maps(segs)
sids, foff, fpos = capi.host_make_reverse(sid, soff, spos, 12)   # both orientations, the executable's own host code
cids, coff, cpos = capi.flatten_maps(contigs)
ctx = capi.Context(0)
ctx.set_maps(0, coff, cpos, capi.host_non_collapse_probs(coff, cpos, 2))
ctx.set_maps(1, foff, fpos, capi.host_non_collapse_probs(foff, fpos, 2))
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
print("checksum", float(np.sum(res["score"].astype(np.float64))))
