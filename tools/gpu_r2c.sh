#!/bin/bash
# round-2 GPU session C: the device-resident alignment loop
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rounds.py tests/test_cli_parity.py -m gpu -x -q --durations=8 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
python - > gpurun_out/r2c_c2_verbose.log 2>&1 <<'PY'
import os, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ampliconreconstructorom_b200 import synth
import cli_cases as cc
tmp = Path(tempfile.mkdtemp())
segs, contigs = synth.make_workload(2000, 2000, 500)
synth.write_cmap(tmp / "s.cmap", segs); synth.write_cmap(tmp / "c.cmap", contigs)
for tag, extra in (("rounds", {}), ("batched", {"SA_NO_ROUNDS": "1"})):
    (tmp / tag).mkdir()
    t0 = time.perf_counter()
    p = subprocess.run([str(cc.NEW_BIN), str(tmp / "s.cmap"), str(tmp / "c.cmap"), f"-prefix={tmp}/{tag}/SA", "-nthreads=16", "-min_labs=12", "-gen=2", "-ngpus=1"],
                       capture_output=True, text=True, env=dict(os.environ, SA_VERBOSE="1", **extra))
    print(tag, "wall", time.perf_counter() - t0, "rc", p.returncode)
    print("\n".join(l for l in p.stderr.splitlines() if "chunk:" not in l or "rounds" in l))
only_a, only_b, diff, nf = cc.compare_dirs(tmp / "rounds", tmp / "batched")
print("rounds vs batched:", nf, "files; only_rounds", only_a[:3], "only_batched", only_b[:3], "different", diff[:3])
PY
tail -12 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_c2_verbose.log
