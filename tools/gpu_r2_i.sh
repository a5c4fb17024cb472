#!/bin/bash
# Round 2, GPU job I: A/B of two-M-tile plain GEMMs for wide outputs (debug build switch ZVX_MT2_N), stage split.
set -x
mkdir -p gpurun_out
ZVX_BUILD_DEBUG=1 python __graft_entry__.py > gpurun_out/build_debug.log 2>&1; tail -1 gpurun_out/build_debug.log
for n in 64 192 2048; do
  echo "ZVX_MT2_N=$n"
  ZVX_MT2_N=$n timeout 300 python tools/bench_configs.py --config 2 2>/dev/null | tail -1 | cut -c1-220
done
ZVX_MT2_N=2048 timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -q -x -k "gemm or decoder or forward_against" 2>&1 | tail -3
