#!/bin/bash
# Round 2, GPU job V: CTA-pair mode of gemm_tc (full-width filter-row convs) — kernel tests, decoder parity, stage split.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python tools/bench_configs.py --config 2 2>&1 | tail -1
