#!/bin/bash
# Round 2, GPU job H: whole GPU suite + bench (both arms) after the voc_pair / wide-conv work.
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
grep "\[parity\]" gpurun_out/pytest_gpu.log > gpurun_out/parity_fullsize.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python tools/bench_configs.py --config 2 > gpurun_out/config2_h.jsonl 2>&1; tail -1 gpurun_out/config2_h.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_h.csv python tools/prof_step.py > gpurun_out/launches_h.log 2>&1
