#!/bin/bash
# Round 2, GPU job F: voc_pair epilogue diet — parity, stage split, launch list, then a debug-build phase trace with three wait hints.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "vocoder or forward_against or full_size" > gpurun_out/pytest_voc_f.log 2>&1; tail -3 gpurun_out/pytest_voc_f.log
timeout 300 python tools/bench_configs.py --config 2 > gpurun_out/config2_f.jsonl 2>&1; tail -1 gpurun_out/config2_f.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_f.csv python tools/prof_step.py > gpurun_out/launches_f.log 2>&1
grep "voc_p" gpurun_out/launches_f.csv | awk -F'","' '{print $5, $(NF)}'
ZVX_BUILD_DEBUG=1 python __graft_entry__.py > gpurun_out/build_debug.log 2>&1; tail -1 gpurun_out/build_debug.log
for ns in 200 2000 20000; do
  ZVX_PAIR_WAIT_NS=$ns timeout 300 python tools/bench_configs.py --config 2 2>/dev/null | tail -1 | cut -c1-200
done
ZVX_VOC_DBG=1 timeout 300 python tools/prof_step.py --warmup 1 2> gpurun_out/voc_dbg_f.txt > /dev/null
grep "voc dbg" gpurun_out/voc_dbg_f.txt | tail -9
