#!/bin/bash
# Round 2, GPU job B: the new voc_pair kernel — vocoder / forward parity, N=1 bench (no CPU leg), stage split, launch list.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "vocoder or forward or full_size or config5" > gpurun_out/pytest_voc.log 2>&1
tail -15 gpurun_out/pytest_voc.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
tail -c 2500 gpurun_out/bench_b.json; tail -3 gpurun_out/bench_b.err
timeout 300 python tools/bench_configs.py --config 2 > gpurun_out/config2_b.jsonl 2>&1; tail -2 gpurun_out/config2_b.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_b.csv python tools/prof_step.py > gpurun_out/launches_b.log 2>&1
grep -c voc_pair gpurun_out/launches_b.csv
