#!/bin/bash
# round-2 GPU session A: new parity tests, throughput microbenchmarks, baseline bench, whole-run phase timings, STORE-kernel profile
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 120 ./tools/ubench12 > gpurun_out/r2a_ubench12.txt 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 400 python tests/perf/c5_cohort.py 1 1.0 c2 > gpurun_out/r2a_c2_whole.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fill_kernel<1" -c 4 \
  -o gpurun_out/prof_r2a_store -f python tests/perf/c5_cohort.py 1 0.3 c2 > gpurun_out/r2a_ncu_store.log 2>&1
tail -5 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_ubench12.txt; cat gpurun_out/r2a_bench.json; tail -30 gpurun_out/r2a_c2_whole.log
