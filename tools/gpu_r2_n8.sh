#!/bin/bash
# Round 2, multi-GPU bench line of the final build (global batch 32 x N on rank 0; sharded forward; both host deliveries).
set -x
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 3000 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
