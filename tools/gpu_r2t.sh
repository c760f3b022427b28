#!/bin/bash
# round-2 GPU session T (1 GPU): packed f32x2 bound chain -- parity (whole gpu suite), scoring rate, C2 whole run; exit-path experiment
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -7 gpurun_out/r2t_pytest.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2t_bench_noextras.json 2> gpurun_out/r2t_bench_noextras.err; tail -2 gpurun_out/r2t_bench_noextras.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2t_bench_noextras.json").readline())
print("value", d["value"], "ms_per_step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e_scoring", d.get("e2e_scoring", {}).get("value"))
PY
python - > gpurun_out/r2t_c2_verbose.log 2>&1 <<'PY'
import os, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ampliconreconstructorom_b200 import synth
import cli_cases as cc
tmp = Path(tempfile.mkdtemp())
segs, contigs = synth.make_workload(2000, 2000, 500)
synth.write_cmap(tmp / "s.cmap", segs); synth.write_cmap(tmp / "c.cmap", contigs)
small_s = {k: segs[k] for k in sorted(segs)[:20]}; small_c = {k: contigs[k] for k in sorted(contigs)[:60]}
synth.write_cmap(tmp / "ss.cmap", small_s); synth.write_cmap(tmp / "sc.cmap", small_c)
for tag, extra in (("one", {}), ("fastexit", {"SA_FAST_EXIT": "1"})):
    (tmp / tag).mkdir()
    t0 = time.perf_counter()
    p = subprocess.run([str(cc.NEW_BIN), str(tmp / "s.cmap"), str(tmp / "c.cmap"), f"-prefix={tmp}/{tag}/SA", "-nthreads=16", "-min_labs=12", "-gen=2", "-ngpus=1"],
                       capture_output=True, text=True, env=dict(os.environ, SA_VERBOSE="1", **extra))
    print(tag, "process wall", round(time.perf_counter() - t0, 3), "rc", p.returncode)
    print("\n".join(l for l in p.stderr.splitlines() if l.startswith("SegAligner")))
    for rep in range(2):   # back-to-back small runs: does skipping the runtime's exit handlers move the cost to the next start?
        t0 = time.perf_counter()
        for k in range(4):
            subprocess.run([str(cc.NEW_BIN), str(tmp / "ss.cmap"), str(tmp / "sc.cmap"), f"-prefix={tmp}/{tag}/small", "-nthreads=4", "-min_labs=12", "-gen=2", "-ngpus=1"],
                           capture_output=True, text=True, env=dict(os.environ, **extra))
        print(tag, "4 small runs back to back:", round(time.perf_counter() - t0, 3), "s")
only_a, only_b, diff, nf = cc.compare_dirs(tmp / "one", tmp / "fastexit")
print("one vs fastexit:", nf, "files; different", diff[:3], only_a[:3], only_b[:3])
PY
cat gpurun_out/r2t_c2_verbose.log
