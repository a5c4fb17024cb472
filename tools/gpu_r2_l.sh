#!/bin/bash
# Round 2, GPU job L: source-level ncu capture of the speaker net's 32-channel 3x3 conv (first gemm_tc launches of a step)
# and of HiFi-GAN's 64-channel stage convs.
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -c 2 \
    -o gpurun_out/gemm_spk32 python tools/prof_step.py > gpurun_out/ncu_gemm_spk32.log 2>&1
tail -2 gpurun_out/ncu_gemm_spk32.log
ncu -i gpurun_out/gemm_spk32.ncu-rep --page raw --csv > gpurun_out/gemm_spk32_raw.csv 2>/dev/null
ls -la gpurun_out/gemm_spk32.ncu-rep
timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -q -x 2>&1 | tail -3
