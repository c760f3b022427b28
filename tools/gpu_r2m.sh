#!/bin/bash
# round-2 GPU session M (1 GPU): coupled-units probe (grid-size threshold of the hand-over stall); digest kernel tests + rate
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python tests/perf/cpl_probe.py > gpurun_out/r2m_cpl_probe.txt 2>&1; echo "rc=$?" >> gpurun_out/r2m_cpl_probe.txt
cat gpurun_out/r2m_cpl_probe.txt


