#!/bin/bash
# round-2 GPU session E: bench at N=1 (full line), ncu launch list of the bench, full captures of the scoring and rounds kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2e_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2e_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fill_kernel<\(int\)0" -c 1 --launch-skip 1 \
   -o gpurun_out/prof_r2e_score -f python bench.py --steps 1 --warmup 1 --no-extras --scale 0.3 > gpurun_out/r2e_ncu_score.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fill_kernel<\(int\)3" -c 2 \
   -o gpurun_out/prof_r2e_rounds -f python tests/perf/c5_cohort.py 1 0.3 c2 > gpurun_out/r2e_ncu_rounds.log 2>&1
cat gpurun_out/r2e_bench_n1.json | head -c 3000; tail -3 gpurun_out/r2e_bench_n1.err; tail -3 gpurun_out/r2e_ncu_score.log; tail -3 gpurun_out/r2e_ncu_rounds.log
