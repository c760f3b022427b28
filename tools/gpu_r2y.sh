#!/bin/bash
# round-2 GPU session Y (1 GPU): rounds stop below the score a file needs -- rounds + CLI parity tests, C2 whole run with and without
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rounds.py tests/test_cli_parity.py -m gpu -x -q --durations=3 > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -12 gpurun_out/r2y_pytest.log
timeout 300 python - > gpurun_out/r2y_c2_verbose.log 2>&1 <<'PY'
import os, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ampliconreconstructorom_b200 import synth
import cli_cases as cc
tmp = Path(tempfile.mkdtemp())
segs, contigs = synth.make_workload(2000, 2000, 500)
synth.write_cmap(tmp / "s.cmap", segs); synth.write_cmap(tmp / "c.cmap", contigs)
for tag, extra in (("stop", {}), ("nostop", {"SA_NO_EARLY_STOP": "1"})):
    (tmp / tag).mkdir()
    t0 = time.perf_counter()
    p = subprocess.run([str(cc.NEW_BIN), str(tmp / "s.cmap"), str(tmp / "c.cmap"), f"-prefix={tmp}/{tag}/SA", "-nthreads=16", "-min_labs=12", "-gen=2", "-ngpus=1"],
                       capture_output=True, text=True, env=dict(os.environ, SA_VERBOSE="1", **extra))
    print(tag, "process wall", round(time.perf_counter() - t0, 3), "rc", p.returncode)
    print(p.stderr[-6000:])
only_a, only_b, diff, nf = cc.compare_dirs(tmp / "stop", tmp / "nostop")
print("stop vs nostop:", nf, "files; different", diff[:3], only_a[:3], only_b[:3])
PY
cat gpurun_out/r2y_c2_verbose.log
