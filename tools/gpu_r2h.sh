#!/bin/bash
# round-2 GPU session H (4 GPUs): CUDA start-up probe -- processes vs threads, concurrent vs serialised
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=persistence_mode --format=csv
for rep in 1 2; do
./tools/ctx_probe procs 1 0
./tools/ctx_probe procs 4 0
./tools/ctx_probe procs 4 1
./tools/ctx_probe procs 4 2
./tools/ctx_probe threads 4 0
./tools/ctx_probe threads 4 1
./tools/ctx_probe procs 2 0
done
} > gpurun_out/r2h_ctx_probe.txt 2>&1
cat gpurun_out/r2h_ctx_probe.txt
