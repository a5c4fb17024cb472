#!/bin/bash
# Round 2, GPU job E: ncu --set full with source counters on voc_pair launches (release build).
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:voc_pair -c 9 \
    -o gpurun_out/voc_pair_full python tools/prof_step.py > gpurun_out/ncu_voc_pair.log 2>&1
tail -3 gpurun_out/ncu_voc_pair.log
ncu -i gpurun_out/voc_pair_full.ncu-rep --page raw --csv > gpurun_out/voc_pair_raw.csv 2>/dev/null
ls -la gpurun_out/voc_pair_full.ncu-rep
