#!/bin/bash
# Round 2, GPU job O: fused attention — parity suite, stage split, launch list.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_attention.py -m gpu -q -x 2>&1 | tail -8
timeout 300 python tools/bench_configs.py --config 2 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_attn.csv python tools/prof_step.py > gpurun_out/launches_attn.log 2>&1
tail -2 gpurun_out/launches_attn.log
grep -c attn_fused gpurun_out/launches_attn.csv
grep attn_fused gpurun_out/launches_attn.csv | head -3
