#!/bin/bash
# round-2 GPU session S (4 GPUs): whole executable on C2 at -ngpus=4,1,2,4 and with no -ngpus (auto), process walls; C5 cohort auto vs 1 GPU
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python - > gpurun_out/r2s_c2_verbose.log 2>&1 <<'PY'
import os, subprocess, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ampliconreconstructorom_b200 import synth
import cli_cases as cc
tmp = Path(tempfile.mkdtemp())
segs, contigs = synth.make_workload(2000, 2000, 500)
synth.write_cmap(tmp / "s.cmap", segs); synth.write_cmap(tmp / "c.cmap", contigs)
for ng in (4, 1, 2, 4, 0):
    tag = f"g{ng}"
    (tmp / tag).mkdir(exist_ok=True)
    t0 = time.perf_counter()
    p = subprocess.run([str(cc.NEW_BIN), str(tmp / "s.cmap"), str(tmp / "c.cmap"), f"-prefix={tmp}/{tag}/SA", "-nthreads=16", "-min_labs=12", "-gen=2"] + ([f"-ngpus={ng}"] if ng else []),
                       capture_output=True, text=True, env=dict(os.environ, SA_VERBOSE="1"))
    print(tag, "(auto)" if not ng else "", "process wall", round(time.perf_counter() - t0, 3), "rc", p.returncode)
    print("\n".join(l for l in p.stderr.splitlines() if "chunk" not in l and "device-resident" not in l and "can reach" not in l))
    time.sleep(1.5)
for other in ("g2", "g4", "g0"):
    only_a, only_b, diff, nf = cc.compare_dirs(tmp / "g1", tmp / other)
    print("g1 vs", other, ":", nf, "files; only_one", only_a[:3], "only_other", only_b[:3], "different", diff[:3])
PY
cat gpurun_out/r2s_c2_verbose.log
timeout 600 python tests/perf/c5_cohort.py 2 1.0 c5 > gpurun_out/r2s_c5_cohort.log 2>&1; grep -v "^  " gpurun_out/r2s_c5_cohort.log | tail -30
