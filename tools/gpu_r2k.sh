#!/bin/bash
# round-2 GPU session K (1 GPU): filtered-rounds test, C3 detection sweep with per-chunk timings, digest ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rounds.py -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
timeout 600 python tests/perf/c3_detection.py 64 1 > gpurun_out/r2k_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:digest_kernel -c 1 -o gpurun_out/prof_r2k_digest -f python tests/perf/digest_probe.py > gpurun_out/r2k_ncu.log 2>&1
tail -5 gpurun_out/r2k_pytest.log; tail -30 gpurun_out/r2k_c3.log
