#!/bin/bash
# Round 2, multi-GPU job: bench.py through torchrun on N GPUs (north_star's scatter -> forward -> gather-v path).
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 4000 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
