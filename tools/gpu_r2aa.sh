#!/bin/bash
# round-2 GPU session AA (1 GPU): memcheck of the rounds kernel's new paths (used strips, end windows), C5 cohort timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --target-processes all python -m pytest tests/test_gpu_rounds.py -m gpu -x -q > gpurun_out/r2aa_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2aa_memcheck.log
tail -8 gpurun_out/r2aa_memcheck.log
timeout 300 python tests/perf/c5_cohort.py 1 > gpurun_out/r2aa_c5.log 2>&1; echo "c5 rc=$?" >> gpurun_out/r2aa_c5.log
cat gpurun_out/r2aa_c5.log | tail -40
