#!/bin/bash
# round-2 GPU session L (1 GPU): coupled-units wavefront -- parity through every shape, detection fill rate legacy vs coupled, C3 sweep
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -x -q --durations=5 > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -8 gpurun_out/r2l_pytest.log
for sh in legacy coupled; do
  echo "== SA_SHAPES=$sh" >> gpurun_out/r2l_quick_detect.txt
  SA_SHAPES=$sh timeout 300 python tests/perf/quick_detect.py 64 >> gpurun_out/r2l_quick_detect.txt 2>&1
done
cat gpurun_out/r2l_quick_detect.txt
timeout 600 python tests/perf/c3_detection.py 64 1 > gpurun_out/r2l_c3.log 2>&1
tail -32 gpurun_out/r2l_c3.log
