"""C5: cohort sweep -- 50 synthetic amplicons x 40 segments against 3,000 genome-wide DLE1 contigs
(~5 x 10^5 labels), the whole standard SegAligner run through the executable, sharded over the GPUs of the box by
problem index (-ngpus=N).  Runs it on 1 GPU and on N GPUs, checks the two output directories are byte-identical and
prints the phase timings.  The reference is far too slow for this size (~20 CPU core-hours); its parity is covered at
CLI level by tests/test_cli_parity.py and tests/test_gpu_fullsize.py.
usage: python tests/perf/c5_cohort.py [n_gpus] [scale] [c5|c2]"""
import os, shutil, subprocess, sys, tempfile, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from ampliconreconstructorom_b200 import synth
import cli_cases as cc

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
tmp = Path(tempfile.mkdtemp(prefix="c5_"))
config = sys.argv[3] if len(sys.argv) > 3 else "c5"
if config == "c2":  # the bench.py workload (2,000 segments x 500 long DLE1 contigs) through the whole executable
    segs, contigs = synth.make_workload(2000, int(2000 * scale), int(500 * scale))
else:
    segs, contigs = synth.make_workload(5, int(2000 * scale), int(3000 * scale), seg_median=40, contig_median=150,
                                        contig_sigma=0.5, contig_clip=(40, 1200))
synth.write_cmap(tmp / "segs.cmap", segs)
synth.write_cmap(tmp / "contigs.cmap", contigs)
nx = sum(len(m) - 1 for m in segs.values()); nb = sum(len(m) - 1 for m in contigs.values())
cells = 2.0 * nx * nb
print(f"{config.upper()}: {len(segs)} segments ({nx} labels) x 2 orientations vs {len(contigs)} contigs ({nb} labels): {cells:.3e} scoring cells")
env = dict(os.environ, SA_VERBOSE="1")
times = {}
for g in sorted({1, n_gpus}):
    out = tmp / f"g{g}"; out.mkdir()
    for rep in range(2):
        t0 = time.perf_counter()
        p = subprocess.run([str(cc.NEW_BIN), str(tmp / "segs.cmap"), str(tmp / "contigs.cmap"), f"-prefix={out}/SA",
                            "-nthreads=8", "-min_labs=12", "-gen=2", f"-ngpus={g}"], capture_output=True, text=True, env=env)
        times[g] = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr[-2000:]
    print(f"--- {g} GPU(s): process wall {times[g]:.2f} s -> {cells/times[g]/1e9:.1f} GCUPS whole process")
    print("\n".join(l for l in p.stderr.splitlines() if "SegAligner(b200)" in l))
if n_gpus > 1:
    only_a, only_b, diff, nf = cc.compare_dirs(tmp / "g1", tmp / f"g{n_gpus}")
    print(f"1 GPU vs {n_gpus} GPUs: {nf} output files, only_1={only_a[:3]} only_N={only_b[:3]} different={diff[:3]}")
    assert not only_a and not only_b and not diff
shutil.rmtree(tmp)
