"""Developer probe: the same C2-shaped problems through the scoring fill (nothing stored) and the alignment fill
(predecessor codes stored + traceback): fill-kernel GCUPS of both."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ampliconreconstructorom_b200 import capi, synth
n_segs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n_contigs = int(sys.argv[2]) if len(sys.argv) > 2 else 100
segs, contigs = synth.make_workload(2000, n_segs, n_contigs)
sid, soff, spos = capi.flatten_maps(segs)
sids, foff, fpos = capi.host_make_reverse(sid, soff, spos, 12)
cids, coff, cpos = capi.flatten_maps(contigs)
ctx = capi.Context(0)
ctx.set_maps(0, coff, cpos, capi.host_non_collapse_probs(coff, cpos, 2))
ctx.set_maps(1, foff, fpos, capi.host_non_collapse_probs(foff, fpos, 2))
cc, ss = np.meshgrid(np.arange(len(cids)), np.arange(len(sids)), indexing="ij")
pr = ctx.problems(cc.ravel(), ss.ravel())
cells = float((ctx.n_labels[0][pr["contig"]] * ctx.n_labels[1][pr["seg"]]).sum())
mode = ctx.mode(0, 0, 0, 6, 0)
for rep in range(2):
    ctx.score_batch(mode, pr); a = ctx.timing().fill_ms
    ctx.align_batch(mode, pr); t = ctx.timing(); b = t.fill_ms
    print(f"rep{rep}: {pr.size} problems {cells:.3e} cells: scoring fill {a:.1f} ms ({cells/a/1e6:.1f} GCUPS), "
          f"alignment fill {b:.1f} ms ({cells/b/1e6:.1f} GCUPS), traceback {t.post_ms:.1f} ms")
