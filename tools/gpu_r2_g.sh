#!/bin/bash
# Round 2, GPU job G: wide two-M-tile FFN conv (single-buffered TMEM) + voc_pair entry prefetch: parity, timing, launch list, trace.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -s > gpurun_out/pytest_g.log 2>&1; tail -3 gpurun_out/pytest_g.log
grep "\[parity\]" gpurun_out/pytest_g.log | head -12
timeout 300 python tools/bench_configs.py --config 2 > gpurun_out/config2_g.jsonl 2>&1; tail -1 gpurun_out/config2_g.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_g.csv python tools/prof_step.py > gpurun_out/launches_g.log 2>&1
grep "voc_p" gpurun_out/launches_g.csv | awk -F'","' '{print $5, $(NF)}'
ZVX_BUILD_DEBUG=1 python __graft_entry__.py > gpurun_out/build_debug.log 2>&1; tail -1 gpurun_out/build_debug.log
ZVX_XR1_WIDE=0 timeout 300 python tools/bench_configs.py --config 2 2>/dev/null | tail -1 | cut -c1-200
ZVX_VOC_DBG=1 timeout 300 python tools/prof_step.py --warmup 1 2> gpurun_out/voc_dbg_g.txt > /dev/null
grep "voc dbg" gpurun_out/voc_dbg_g.txt | tail -9
