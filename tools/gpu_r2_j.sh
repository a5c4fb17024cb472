#!/bin/bash
# Round 2, GPU job J: two CTAs per SM for narrow-output gemm_tc launches: parity, then A/B of the stage split (debug build switch).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py tests/test_gpu_frontend.py -m gpu -q -x > gpurun_out/pytest_j.log 2>&1; tail -3 gpurun_out/pytest_j.log
timeout 300 python tools/bench_configs.py --config 2 > gpurun_out/config2_j.jsonl 2>&1; tail -1 gpurun_out/config2_j.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_j.csv python tools/prof_step.py > gpurun_out/launches_j.log 2>&1
ZVX_BUILD_DEBUG=1 python __graft_entry__.py > gpurun_out/build_debug.log 2>&1; tail -1 gpurun_out/build_debug.log
for v in 0 1; do echo "ZVX_TC_TWO=$v"; ZVX_TC_TWO=$v timeout 300 python tools/bench_configs.py --config 2 2>/dev/null | tail -1 | cut -c1-220; done
