#!/bin/bash
# Round 2, GPU job K: quick parity + timing after small kernel changes.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frontend.py -m gpu -q -x > gpurun_out/pytest_k.log 2>&1; tail -3 gpurun_out/pytest_k.log
timeout 300 python tools/bench_configs.py --config 2 > gpurun_out/config2_k.jsonl 2>&1; tail -1 gpurun_out/config2_k.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_k.csv python tools/prof_step.py > gpurun_out/launches_k.log 2>&1
grep -E "skinny|conv_post|hw_sum" gpurun_out/launches_k.csv | awk -F'","' '{print substr($5,1,50), $(NF)}' | sort | uniq -c | head
