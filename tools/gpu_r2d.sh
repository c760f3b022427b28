#!/bin/bash
# round-2 GPU session D: full gpu suite on the SPMD engine + skip; C2 whole run, 1 process and 2 oversubscribed processes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
python - > gpurun_out/r2d_c2_verbose.log 2>&1 <<'PY'
import os, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ampliconreconstructorom_b200 import synth
import cli_cases as cc
tmp = Path(tempfile.mkdtemp())
segs, contigs = synth.make_workload(2000, 2000, 500)
synth.write_cmap(tmp / "s.cmap", segs); synth.write_cmap(tmp / "c.cmap", contigs)
for tag, extra, ng in (("one", {}, 1), ("noskip", {"SA_NO_SKIP": "1"}, 1), ("two", {"SA_GPU_OVERSUBSCRIBE": "1"}, 2)):
    (tmp / tag).mkdir()
    t0 = time.perf_counter()
    p = subprocess.run([str(cc.NEW_BIN), str(tmp / "s.cmap"), str(tmp / "c.cmap"), f"-prefix={tmp}/{tag}/SA", "-nthreads=16", "-min_labs=12", "-gen=2", f"-ngpus={ng}"],
                       capture_output=True, text=True, env=dict(os.environ, SA_VERBOSE="1", **extra))
    print(tag, "wall", time.perf_counter() - t0, "rc", p.returncode)
    print(p.stderr)
for other in ("noskip", "two"):
    only_a, only_b, diff, nf = cc.compare_dirs(tmp / "one", tmp / other)
    print("one vs", other, ":", nf, "files; only_one", only_a[:3], "only_other", only_b[:3], "different", diff[:3])
PY
tail -12 gpurun_out/r2d_pytest.log; cat gpurun_out/r2d_c2_verbose.log
