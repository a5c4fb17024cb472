#!/bin/bash
# Round 2, GPU job D: voc_pair with the MRF accumulator folded in during the last MMA phase; MMA cost vs operand K-half distance.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "vocoder or forward_against" > gpurun_out/pytest_voc_d.log 2>&1; tail -3 gpurun_out/pytest_voc_d.log
timeout 300 python tools/bench_configs.py --config 2 > gpurun_out/config2_d.jsonl 2>&1; tail -1 gpurun_out/config2_d.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_d.csv python tools/prof_step.py > gpurun_out/launches_d.log 2>&1
grep "voc_p" gpurun_out/launches_d.csv | awk -F'","' '{print $5, $(NF)}'
timeout 120 ./build/ubench_mma --lbo > gpurun_out/ubench_lbo.txt 2>&1; cat gpurun_out/ubench_lbo.txt
