"""C3: hg19 BspQI detection sweep.  64 synthetic unaligned regions (21-500 labels, seed 3) cut from
the real hg19 BspQI reference CMAP and passed through the optical-map error model, searched against the whole
reference in -detection mode (the ARAlignDetect.py:408-411 call).  Times the B200 executable on the full sweep and
the unmodified reference on a subset (byte-compared), and reports both as GCUPS."""
import gzip, os, shutil, subprocess, sys, tempfile, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from ampliconreconstructorom_b200 import capi, synth
import cli_cases as cc

data = ROOT / "oracle" / "_ref" / "data"
n_regions = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_ref_regions = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ngpus = sys.argv[3] if len(sys.argv) > 3 else "1"
tmp = Path(tempfile.mkdtemp(prefix="c3_"))
hg = tmp / "hg19.cmap"
with gzip.open(data / "hg19_BspQI.cmap.gz", "rb") as f, open(hg, "wb") as g:
    shutil.copyfileobj(f, g)
ids, off, pos = capi.host_parse_cmap(hg)
labels = np.diff(off) - 1
rng = np.random.default_rng(3)
regions = {}
big = [k for k in range(len(ids)) if labels[k] > 2000]
for r in range(n_regions):
    k = big[int(rng.integers(0, len(big)))]
    n = int(rng.integers(21, 501))
    chrom = pos[off[k]:off[k + 1] - 1].astype(np.float64)
    s = int(rng.integers(1, chrom.size - n - 2))
    regions[r + 1] = synth.noisy_contig(rng, chrom, s, n)
total_nx = sum(len(m) - 1 for m in regions.values())
cells = 2.0 * total_nx * float(labels.sum())
synth.write_cmap(tmp / "regions.cmap", regions)
flags = ["-nthreads=%d" % (os.cpu_count() or 1), "-min_labs=10", "-gen=1", "-detection", f"-ngpus={ngpus}"]
env = dict(os.environ, SA_VERBOSE="1")
for rep in range(2):
    t0 = time.perf_counter()
    p = subprocess.run([str(cc.NEW_BIN), str(tmp / "regions.cmap"), str(hg), f"-prefix={tmp}/new_SA"] + flags,
                       capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    assert p.returncode == 0, p.stderr
print(p.stderr)
n_aln = len([f for f in tmp.iterdir() if f.name.startswith("new_SA") and f.name.endswith("_aln.txt")])
print(f"B200: {n_regions} regions, {total_nx} labels x2 orientations vs hg19 ({int(labels.sum())} labels): {cells:.3e} cells, "
      f"process wall {dt:.2f} s -> {cells/dt/1e9:.2f} GCUPS whole process; {n_aln} alignment files")
# reference on a subset, byte-compared with ours on the same subset
sub = {k: regions[k] for k in sorted(regions)[:n_ref_regions]}
synth.write_cmap(tmp / "sub.cmap", sub)
sub_cells = 2.0 * sum(len(m) - 1 for m in sub.values()) * float(labels.sum())
(tmp / "r").mkdir(); (tmp / "n").mkdir()
t0 = time.perf_counter()
subprocess.run([str(cc.REF_BIN), str(tmp / "sub.cmap"), str(hg), f"-prefix={tmp}/r/SA"] + flags[:-1], stdout=subprocess.DEVNULL, check=True)
dtr = time.perf_counter() - t0
subprocess.run([str(cc.NEW_BIN), str(tmp / "sub.cmap"), str(hg), f"-prefix={tmp}/n/SA"] + flags, stdout=subprocess.DEVNULL, check=True)
only_a, only_b, diff, nf = cc.compare_dirs(tmp / "r", tmp / "n")
print(f"reference: {len(sub)} regions, {sub_cells:.3e} cells, wall {dtr:.1f} s on {os.cpu_count()} threads -> {sub_cells/dtr/1e9:.4f} GCUPS; "
      f"outputs vs ours: {nf} files, only_ref={only_a} only_ours={only_b} different={diff}")
print(f"whole-process speed-up on the sweep (reference extrapolated linearly in cells): {(cells/dt)/(sub_cells/dtr):.0f}x")
shutil.rmtree(tmp)
