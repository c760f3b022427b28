#!/bin/bash
# round-2 GPU session B: full gpu test suite, new bench.py, tip-phase round statistics, STORE-kernel profile
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?" >> gpurun_out/r2b_bench.err
python - > gpurun_out/r2b_c2_verbose.log 2>&1 <<'PY'
import os, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ampliconreconstructorom_b200 import synth
import cli_cases as cc
tmp = Path(tempfile.mkdtemp())
segs, contigs = synth.make_workload(2000, 2000, 500)
synth.write_cmap(tmp / "s.cmap", segs); synth.write_cmap(tmp / "c.cmap", contigs)
t0 = time.perf_counter()
p = subprocess.run([str(cc.NEW_BIN), str(tmp / "s.cmap"), str(tmp / "c.cmap"), f"-prefix={tmp}/SA", "-nthreads=16", "-min_labs=12", "-gen=2", "-ngpus=1"],
                   capture_output=True, text=True, env=dict(os.environ, SA_VERBOSE="1"))
print("wall", time.perf_counter() - t0, "rc", p.returncode)
print(p.stderr)
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"fill_kernel<\(int\)1, \(bool\)0, \(int\)1," -c 3 -o gpurun_out/prof_r2b_store -f \
  python tests/perf/c5_cohort.py 1 0.3 c2 > gpurun_out/r2b_ncu_store.log 2>&1
tail -15 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err; tail -5 gpurun_out/r2b_ncu_store.log
