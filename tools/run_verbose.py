"""Developer probe: one whole-executable run of a synthetic workload (c2 | c5) on one GPU with SA_VERBOSE=1, all of stderr shown.
usage: python tools/run_verbose.py [c2|c5] [scale] [NGPUS=n] [ENV=VALUE ...]"""
import os, subprocess, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from ampliconreconstructorom_b200 import synth
import cli_cases as cc
config = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
extra = dict(a.split("=", 1) for a in sys.argv[3:])
ngpus = extra.pop("NGPUS", "1")  # NGPUS=4: one process per GPU (-ngpus=4)
tmp = Path(tempfile.mkdtemp(prefix="rv_"))
if config == "c2":
    segs, contigs = synth.make_workload(2000, int(2000 * scale), int(500 * scale))
else:
    segs, contigs = synth.make_workload(5, int(2000 * scale), int(3000 * scale), seg_median=40, contig_median=150,
                                        contig_sigma=0.5, contig_clip=(40, 1200))
synth.write_cmap(tmp / "s.cmap", segs); synth.write_cmap(tmp / "c.cmap", contigs)
(tmp / "out").mkdir()
t0 = time.perf_counter()
p = subprocess.run([str(cc.NEW_BIN), str(tmp / "s.cmap"), str(tmp / "c.cmap"), f"-prefix={tmp}/out/SA", "-nthreads=16", "-min_labs=12", "-gen=2", f"-ngpus={ngpus}"],
                   capture_output=True, text=True, env=dict(os.environ, SA_VERBOSE="1", **extra))
print(config, scale, extra, "ngpus", ngpus, "process wall", round(time.perf_counter() - t0, 3), "rc", p.returncode, "files", len(list((tmp / "out").iterdir())))
print(p.stderr[-8000:])
