#!/bin/bash
# Round 2, GPU job R: state of the fused attention build — kernel tests, parity suites, stand-alone timing, stage split, bench line.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention.py -m gpu -q -x 2>&1 | tail -4
timeout 200 python tools/attn_bench.py 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_configs.py --config 2 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json | head -c 600; head -c 700 gpurun_out/bench.json
