#!/bin/bash
# A/B: two M tiles per CTA tile in the filter-row tap-reuse path (ZVX_XR1_MT)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -q -k "conv1d or full_size or golden or hifigan or vocoder" > gpurun_out/pytest_xr1mt.log 2>&1
tail -3 gpurun_out/pytest_xr1mt.log
rm -f gpurun_out/ab_xr1_mt.jsonl
for v in 1 2; do
  echo "== ZVX_XR1_MT=$v" | tee -a gpurun_out/ab_xr1_mt.jsonl
  ZVX_XR1_MT=$v timeout 200 python tools/bench_configs.py --config 2 --iters 15 2>/dev/null | tee -a gpurun_out/ab_xr1_mt.jsonl
done
