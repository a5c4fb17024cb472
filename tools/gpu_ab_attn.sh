#!/bin/bash
# A/B: attention batch slicing (ZVX_ATTN_SLICE_BYTES), configs[1] stage split + decoder parity under slicing
mkdir -p gpurun_out
for sb in 0 100663296 50331648 25165824; do
  echo "== ZVX_ATTN_SLICE_BYTES=$sb"
  ZVX_ATTN_SLICE_BYTES=$sb timeout 200 python tools/bench_configs.py --config 2 --iters 15 2>/dev/null | tee -a gpurun_out/ab_attn.jsonl
done
ZVX_ATTN_SLICE_BYTES=50331648 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_attn_slice.log 2>&1
tail -3 gpurun_out/pytest_attn_slice.log
