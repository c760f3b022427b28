#!/bin/bash
# round-2 GPU session X (1 GPU): tip-pass statistics on the C2 workload (used labels per contig, kernel time per chunk)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/r2x_c2_verbose.log 2>&1 <<'PY'
import os, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ampliconreconstructorom_b200 import synth
import cli_cases as cc
tmp = Path(tempfile.mkdtemp())
segs, contigs = synth.make_workload(2000, 2000, 500)
synth.write_cmap(tmp / "s.cmap", segs); synth.write_cmap(tmp / "c.cmap", contigs)
for tag in ("one",):
    (tmp / tag).mkdir()
    t0 = time.perf_counter()
    p = subprocess.run([str(cc.NEW_BIN), str(tmp / "s.cmap"), str(tmp / "c.cmap"), f"-prefix={tmp}/{tag}/SA", "-nthreads=16", "-min_labs=12", "-gen=2", "-ngpus=1"],
                       capture_output=True, text=True, env=dict(os.environ, SA_VERBOSE="1"))
    print(tag, "process wall", round(time.perf_counter() - t0, 3), "rc", p.returncode)
    print(p.stderr[-6000:])
PY
cat gpurun_out/r2x_c2_verbose.log
