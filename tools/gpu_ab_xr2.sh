#!/bin/bash
# A/B: two M tiles per CTA tile in the 2-D tap-reuse path (ZVX_XR2_MT, ZVX_XR2_MT_MAXN)
mkdir -p gpurun_out
ZVX_XR2=2 ZVX_XR2_MAXN=256 timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -k "conv2d" -s > gpurun_out/pytest_xr2_gemm.log 2>&1
tail -2 gpurun_out/pytest_xr2_gemm.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm.py -m gpu -q -k "speaker or full_size or golden or conv2d" > gpurun_out/pytest_xr2_parity.log 2>&1
tail -2 gpurun_out/pytest_xr2_parity.log
for v in "1 128" "2 32" "2 64" "2 128"; do
  set -- $v
  echo "== ZVX_XR2_MT=$1 ZVX_XR2_MT_MAXN=$2" | tee -a gpurun_out/ab_xr2_mt.jsonl
  ZVX_XR2_MT=$1 ZVX_XR2_MT_MAXN=$2 timeout 200 python tools/bench_configs.py --config 2 --iters 15 2>/dev/null | tee -a gpurun_out/ab_xr2_mt.jsonl
done
