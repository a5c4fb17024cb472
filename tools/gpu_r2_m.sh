#!/bin/bash
# Round 2, GPU job M: elect.sync single-thread roles (no per-instruction ELECT / BRA.U.ANY wrapper) — parity + step time.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py tests/test_gpu_vocoder.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python tools/bench_configs.py --config 2 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_elect.json 2> gpurun_out/bench_n1_elect.err
cat gpurun_out/bench_n1_elect.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_elect.csv python tools/prof_step.py > gpurun_out/launches_elect.log 2>&1
tail -2 gpurun_out/launches_elect.log
