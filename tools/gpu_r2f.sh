#!/bin/bash
# round-2 GPU session F (4 GPUs): strong-scaling bench lines at N=2 and N=4, C5 cohort sweep through the executable on 1 vs 4 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2f_smi.txt; nproc >> gpurun_out/r2f_smi.txt
for N in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 \
     > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err; echo "bench N=$N rc=$?" >> gpurun_out/r2f_bench_n$N.err
done
timeout 600 python tests/perf/c5_cohort.py 4 1.0 c5 > gpurun_out/r2f_c5_cohort.log 2>&1
for N in 2 4; do head -c 1500 gpurun_out/r2f_bench_n$N.json; echo; tail -3 gpurun_out/r2f_bench_n$N.err; done; tail -40 gpurun_out/r2f_c5_cohort.log
