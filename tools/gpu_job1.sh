#!/bin/bash
# Round-1 GPU job: parity tests, bench, ncu launch list, ncu --set full of the two dominant kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/prof_step.py > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:voc_pair_kernel \
    -o gpurun_out/voc_pair python tools/prof_step.py > gpurun_out/ncu_voc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel \
    -s 31 -c 9 -o gpurun_out/gemm_tc_dec python tools/prof_step.py > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
