#!/bin/bash
# round-2 GPU session J (1 GPU): full gpu suite after the filtered rounds / digest v2 / SPMD changes; digest probe; C2 whole run with timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
timeout 300 python tests/perf/digest_probe.py > gpurun_out/r2j_digest_probe.txt 2>&1
python - > gpurun_out/r2j_c2_verbose.log 2>&1 <<'PY'
import os, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ampliconreconstructorom_b200 import synth
import cli_cases as cc
tmp = Path(tempfile.mkdtemp())
segs, contigs = synth.make_workload(2000, 2000, 500)
synth.write_cmap(tmp / "s.cmap", segs); synth.write_cmap(tmp / "c.cmap", contigs)
for tag, extra, ng in (("one", {}, 1), ("one_b", {}, 1), ("two_over", {"SA_GPU_OVERSUBSCRIBE": "1"}, 2)):
    (tmp / tag).mkdir()
    t0 = time.perf_counter()
    p = subprocess.run([str(cc.NEW_BIN), str(tmp / "s.cmap"), str(tmp / "c.cmap"), f"-prefix={tmp}/{tag}/SA", "-nthreads=16", "-min_labs=12", "-gen=2", f"-ngpus={ng}"],
                       capture_output=True, text=True, env=dict(os.environ, SA_VERBOSE="1", **extra))
    print(tag, "wall", time.perf_counter() - t0, "rc", p.returncode)
    print(p.stderr)
only_a, only_b, diff, nf = cc.compare_dirs(tmp / "one", tmp / "two_over")
print("one vs two_over:", nf, "files; only_one", only_a[:3], "only_other", only_b[:3], "different", diff[:3])
PY
tail -12 gpurun_out/r2j_pytest.log; cat gpurun_out/r2j_digest_probe.txt; cat gpurun_out/r2j_c2_verbose.log
