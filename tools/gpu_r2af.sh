#!/bin/bash
# round-2 GPU session AF (1 GPU): final numbers after the tip-pass cuts -- gpu suite, bench N=1 (full line), reference arm,
# ncu launch list of the bench command, ncu full capture of the rounds kernel (standard + tip pass of the C2 executable at scale 0.3)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2af_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2af_pytest.log; tail -3 gpurun_out/r2af_pytest.log
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/r2af_bench_n1.json 2> gpurun_out/r2af_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2af_bench_n1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2af_bench_reference.json 2> gpurun_out/r2af_bench_reference.err; echo "ref rc=$?" >> gpurun_out/r2af_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2af_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2af_ncu_launches.log 2>&1
python - <<'PY'
import sys, tempfile
from pathlib import Path
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ampliconreconstructorom_b200 import synth
segs, contigs = synth.make_workload(2000, 600, 150)
d = Path("/tmp/c2s"); d.mkdir(exist_ok=True); (d / "out").mkdir(exist_ok=True)
synth.write_cmap(d / "s.cmap", segs); synth.write_cmap(d / "c.cmap", contigs)
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fill_kernel<\(int\)3" -c 2 \
   -o gpurun_out/prof_r2af_rounds -f ampliconreconstructorom_b200/bin/SegAligner /tmp/c2s/s.cmap /tmp/c2s/c.cmap -prefix=/tmp/c2s/out/SA -nthreads=16 -min_labs=12 -gen=2 -ngpus=1 > gpurun_out/r2af_ncu_rounds.log 2>&1
head -c 2500 gpurun_out/r2af_bench_n1.json; echo; tail -3 gpurun_out/r2af_bench_n1.err; head -c 1200 gpurun_out/r2af_bench_reference.json; echo; tail -2 gpurun_out/r2af_ncu_rounds.log
