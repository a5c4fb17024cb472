#!/bin/bash
# A/B: filter-row tap reuse for narrow 1-D 3-tap convs (ZVX_XR1_K3)
mkdir -p gpurun_out
ZVX_XR1_K3=1 timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -q -k "conv1d or full_size or golden or hifigan or vocoder" > gpurun_out/pytest_xr1k3.log 2>&1
tail -3 gpurun_out/pytest_xr1k3.log
rm -f gpurun_out/ab_xr1_k3.jsonl
for v in 0 1; do
  echo "== ZVX_XR1_K3=$v" | tee -a gpurun_out/ab_xr1_k3.jsonl
  ZVX_XR1_K3=$v timeout 200 python tools/bench_configs.py --config 2 --iters 15 2>/dev/null | tee -a gpurun_out/ab_xr1_k3.jsonl
  ZVX_XR1_K3=$v timeout 200 python tools/bench_configs.py --config 2 --iters 15 --decoder styletts 2>/dev/null | tee -a gpurun_out/ab_xr1_k3.jsonl
done
