#!/bin/bash
# round-2 GPU session R (1 GPU): full gpu suite after the coupled-unit lag, the auto GPU count and the background teardown; detection fill; C3
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=4 > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -9 gpurun_out/r2r_pytest.log
rm -f gpurun_out/r2r_quick_detect.txt
for sh in legacy coupled; do
  echo "== SA_SHAPES=$sh" >> gpurun_out/r2r_quick_detect.txt
  SA_SHAPES=$sh timeout 200 python tests/perf/quick_detect.py 64 >> gpurun_out/r2r_quick_detect.txt 2>&1
done
cat gpurun_out/r2r_quick_detect.txt
timeout 400 python tests/perf/c3_detection.py 64 1 > gpurun_out/r2r_c3.log 2>&1
grep -v "alignment round" gpurun_out/r2r_c3.log | tail -22
