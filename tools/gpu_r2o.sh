#!/bin/bash
# round-2 GPU session O (1 GPU): coupled-unit planning sweep on the detection fill (coupling threshold, unit size, unit cap)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/r2o_sweep.txt
for cfg in "0.25 0.5 64" "0.1 0.5 64" "0.5 0.5 64" "1.0 0.5 64" "0.25 0.25 64" "0.25 1.0 64" "0.5 0.25 64" "0.25 0.5 16" "0.25 0.5 32" "0.05 0.5 64"; do
  set -- $cfg
  echo "== SA_CPL_ALPHA=$1 SA_CPL_UNIT=$2 SA_CPL_UMAX=$3" >> gpurun_out/r2o_sweep.txt
  SA_CPL_ALPHA=$1 SA_CPL_UNIT=$2 SA_CPL_UMAX=$3 timeout 200 python tests/perf/quick_detect.py 64 2>&1 | tail -1 >> gpurun_out/r2o_sweep.txt
done
cat gpurun_out/r2o_sweep.txt
