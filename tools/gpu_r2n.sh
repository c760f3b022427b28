#!/bin/bash
# round-2 GPU session N (1 GPU): full gpu suite with coupled units as the default for large problems; detection fill legacy vs coupled; C3; digest
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -12 gpurun_out/r2n_pytest.log
rm -f gpurun_out/r2n_quick_detect.txt
for sh in legacy coupled; do
  echo "== SA_SHAPES=$sh" >> gpurun_out/r2n_quick_detect.txt
  SA_SHAPES=$sh timeout 200 python tests/perf/quick_detect.py 64 >> gpurun_out/r2n_quick_detect.txt 2>&1
done
cat gpurun_out/r2n_quick_detect.txt
timeout 400 python tests/perf/c3_detection.py 64 1 > gpurun_out/r2n_c3.log 2>&1
tail -32 gpurun_out/r2n_c3.log
timeout 120 python tests/perf/digest_probe.py > gpurun_out/r2n_digest_probe.txt 2>&1; cat gpurun_out/r2n_digest_probe.txt
