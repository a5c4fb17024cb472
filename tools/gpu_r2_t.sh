#!/bin/bash
# Round 2, GPU job T: attention kernel variants — kernel tests (bounded), stand-alone timing of both variants with / without the
# device-built tile list.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_attention.py -m gpu -q -k "${1:-attention}" 2>&1 | tail -15
for v in 1 2; do for pl in 0 2; do timeout 60 python tools/attn_bench.py --variant $v --plan $pl 2>&1 | tail -1; done; done
timeout 300 python tools/bench_configs.py --config 2 2>&1 | tail -1
