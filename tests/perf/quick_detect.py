"""Developer probe: detection-mode scoring (fill STORE+SWAP + harvest) of N synthetic regions vs hg19 via the C ABI."""
import gzip, shutil, sys, tempfile, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from ampliconreconstructorom_b200 import capi, synth
n_regions = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tmp = Path(tempfile.mkdtemp(prefix="qd_"))
with gzip.open(ROOT / "oracle/_ref/data/hg19_BspQI.cmap.gz", "rb") as f, open(tmp / "hg19.cmap", "wb") as g:
    shutil.copyfileobj(f, g)
ids, off, pos = capi.host_parse_cmap(tmp / "hg19.cmap")
labels = np.diff(off) - 1
rng = np.random.default_rng(3)
regions = {}
big = [k for k in range(len(ids)) if labels[k] > 2000]
for r in range(n_regions):
    k = big[int(rng.integers(0, len(big)))]; n = int(rng.integers(21, 501))
    chrom = pos[off[k]:off[k + 1] - 1].astype(np.float64); s = int(rng.integers(1, chrom.size - n - 2))
    regions[r + 1] = synth.noisy_contig(rng, chrom, s, n)
sid, soff, spos = capi.flatten_maps(regions)
fid, foff, fpos = capi.host_make_reverse(sid, soff, spos, 10)
ctx = capi.Context(0)
ctx.set_maps(0, off, pos, capi.host_non_collapse_probs(off, pos, 1))
ctx.set_maps(1, foff, fpos, capi.host_non_collapse_probs(foff, fpos, 1))
cc, ss = np.meshgrid(np.arange(len(ids)), np.arange(len(fid)), indexing="ij")
pr = ctx.problems(cc.ravel(), ss.ravel())
tot = int((labels + 1).sum())
want = (np.rint(np.float32(500) * (labels[pr["contig"]] + 1).astype(np.float32) / np.float32(tot)).astype(np.int32) + 1)
cells = float(labels.sum()) * float((np.diff(foff) - 1).sum())
mode = ctx.mode(0, 1, 0, 6, 1)
for rep in range(2):
    t0 = time.time(); sc, o, got = ctx.detect_batch(mode, pr, want); t1 = time.time()
    tm = ctx.timing()
    import hashlib
    h = hashlib.sha256(sc.tobytes() + got.tobytes()).hexdigest()[:16]
    print(f"rep{rep}: [{h}] {pr.size} problems {cells:.3e} cells: fill {tm.fill_ms:.1f} ms ({cells/tm.fill_ms/1e6:.2f} GCUPS), harvest {tm.post_ms:.1f} ms, call {t1-t0:.3f} s")
shutil.rmtree(tmp)
