#!/bin/bash
# round-2 GPU session I (1 GPU): packed 2-bit digest kernel -- parity tests, throughput probe, ncu capture
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_digest.py -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
timeout 300 python tests/perf/digest_probe.py > gpurun_out/r2i_digest_probe.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:digest_kernel -c 2 -o gpurun_out/prof_r2i_digest -f python tests/perf/digest_probe.py > gpurun_out/r2i_ncu.log 2>&1
tail -5 gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_digest_probe.txt; tail -3 gpurun_out/r2i_ncu.log
