#!/bin/bash
# Round 2, multi-GPU job: shared host windows (every rank delivers over its own PCIe link) next to the rank-0 funnel.
set -x
N=${1:-2}
mkdir -p gpurun_out
df -h /dev/shm | cat
timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --e2e-groups-extra="${2:-2}" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 5000 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
