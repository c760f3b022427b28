#!/bin/bash
# round-2 GPU session P (1 GPU): ncu of the detection fill (coupled units + bulk, STORE_S + SWAP instantiations)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -c 2 -o gpurun_out/prof_r2p_detect_fill -f python tests/perf/quick_detect.py 64 > gpurun_out/r2p_ncu.log 2>&1
tail -5 gpurun_out/r2p_ncu.log
