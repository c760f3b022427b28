"""Developer probe: the executable with SA_VERBOSE=1 on a C5- or C2-shaped workload: per-phase, per-round and per-chunk timings.
usage: python tools/c5half.py [n_gpus] [scale] [c5|c2]"""
import sys, os, subprocess, time, tempfile
from pathlib import Path
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from ampliconreconstructorom_b200 import synth
import cli_cases as cc
tmp = Path(tempfile.mkdtemp())
sc = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
if len(sys.argv) > 3 and sys.argv[3] == "c2":
    segs, contigs = synth.make_workload(2000, int(2000*sc), int(500*sc))
else:
    segs, contigs = synth.make_workload(5, int(2000*sc), int(3000*sc), seg_median=40, contig_median=150, contig_sigma=0.5, contig_clip=(40, 1200))
synth.write_cmap(tmp / "s.cmap", segs); synth.write_cmap(tmp / "c.cmap", contigs)
env = dict(os.environ, SA_VERBOSE="1")
p = subprocess.run([str(cc.NEW_BIN), str(tmp/"s.cmap"), str(tmp/"c.cmap"), f"-prefix={tmp}/SA", "-nthreads=8", "-min_labs=12", "-gen=2", f"-ngpus={sys.argv[1] if len(sys.argv) > 1 else 1}"], capture_output=True, text=True, env=env)
lines = p.stderr.splitlines()
print("\n".join(l for l in lines if "round" in l or "SegAligner" in l or "chunk" in l))
