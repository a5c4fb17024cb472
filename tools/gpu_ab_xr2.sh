#!/bin/bash
# A/B: 2-D tap reuse (ZVX_XR2) in the TMA implicit-GEMM conv: kernel-level parity forced on every conv2d test shape, speaker-net
# parity, configs[1] stage split with / without it and for different N limits
mkdir -p gpurun_out
ZVX_XR2=2 ZVX_XR2_MAXN=256 timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -k "conv2d_3x3" -s > gpurun_out/pytest_xr2_gemm.log 2>&1
tail -2 gpurun_out/pytest_xr2_gemm.log
ZVX_XR2=1 ZVX_XR2_MAXN=256 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "speaker or full_size or golden" > gpurun_out/pytest_xr2_parity.log 2>&1
tail -2 gpurun_out/pytest_xr2_parity.log
for v in "0 128" "1 32" "1 64" "1 128" "1 256"; do
  set -- $v
  echo "== ZVX_XR2=$1 ZVX_XR2_MAXN=$2" | tee -a gpurun_out/ab_xr2.jsonl
  ZVX_XR2=$1 ZVX_XR2_MAXN=$2 timeout 200 python tools/bench_configs.py --config 2 --iters 15 2>/dev/null | tee -a gpurun_out/ab_xr2.jsonl
done
