#!/bin/bash
# round-2 GPU session V (4 GPUs): final strong-scaling bench lines at N=2 and N=4 (torchrun, one C2 batch sharded by problem index)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for N in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 \
     > gpurun_out/r2v_bench_n$N.json 2> gpurun_out/r2v_bench_n$N.err; echo "bench N=$N rc=$?" >> gpurun_out/r2v_bench_n$N.err
done
for N in 2 4; do head -c 1200 gpurun_out/r2v_bench_n$N.json; echo; tail -2 gpurun_out/r2v_bench_n$N.err; done
