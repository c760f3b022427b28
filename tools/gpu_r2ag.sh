#!/bin/bash
# round-2 GPU session AG (2 GPUs): strong-scaling bench line at N=2 after the tip-pass cuts (torchrun, one C2 batch sharded by problem index)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 \
   > gpurun_out/r2ag_bench_n$N.json 2> gpurun_out/r2ag_bench_n$N.err; echo "bench N=$N rc=$?" >> gpurun_out/r2ag_bench_n$N.err
head -c 600 gpurun_out/r2ag_bench_n$N.json; echo; tail -2 gpurun_out/r2ag_bench_n$N.err
