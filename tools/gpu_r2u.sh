#!/bin/bash
# round-2 GPU session U (1 GPU): final numbers -- gpu suite, bench N=1 (full line), reference arm, ncu launch list, ncu full capture of the scoring kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log; tail -3 gpurun_out/r2u_pytest.log
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2u_bench_n1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2u_bench_reference.json 2> gpurun_out/r2u_bench_reference.err; echo "ref rc=$?" >> gpurun_out/r2u_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2u_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fill_kernel<\(int\)0" -c 1 --launch-skip 1 \
   -o gpurun_out/prof_r2u_score -f python bench.py --steps 1 --warmup 1 --no-extras --scale 0.3 > gpurun_out/r2u_ncu_score.log 2>&1
head -c 2500 gpurun_out/r2u_bench_n1.json; echo; tail -3 gpurun_out/r2u_bench_n1.err; head -c 1200 gpurun_out/r2u_bench_reference.json; echo; tail -2 gpurun_out/r2u_ncu_score.log
