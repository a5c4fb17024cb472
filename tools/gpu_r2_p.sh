#!/bin/bash
# Round 2, GPU job P: ncu --set full of the fused attention kernel (one decoder-sized launch) + source page.
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:attn_fused_kernel -c 1 \
    -o gpurun_out/attn_fused python tools/prof_step.py > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
ncu -i gpurun_out/attn_fused.ncu-rep --page raw --csv > gpurun_out/attn_fused_raw.csv 2>/dev/null
ncu -i gpurun_out/attn_fused.ncu-rep --page source --csv > gpurun_out/attn_fused_source.csv 2>/dev/null
ls -la gpurun_out/attn_fused*
