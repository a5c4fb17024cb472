#!/bin/bash
# Round 2, GPU job U (debug build): CTA-pair mode of gemm_tc's plain tiles — kernel tests, A/B on the hot shapes, parity, step.
set -x
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -3
for pr in 0 1; do echo "== pair $pr"; ZVX_GEMM_PAIR=$pr timeout 300 python tools/bench_gemm.py 2>&1 | grep "TF/s" | head -7; ZVX_GEMM_PAIR=$pr timeout 300 python tools/bench_configs.py --config 2 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
