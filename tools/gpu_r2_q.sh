#!/bin/bash
# Round 2, GPU job Q: fused attention iteration — kernel tests, stage split, launch durations of the attention kernel.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/bench_configs.py --config 2 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:attn_fused --csv \
    --log-file gpurun_out/launches_attn.csv python tools/prof_step.py > gpurun_out/launches_attn.log 2>&1
grep attn_fused gpurun_out/launches_attn.csv | awk -F'","' '{print $NF}' | head -6
