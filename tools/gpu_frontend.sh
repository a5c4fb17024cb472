#!/bin/bash
# Front-end GPU job: parity tests of csrc/frontend.cu, its measurement, and one ncu --set full capture of the mel kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontend.py -m gpu -q -s > gpurun_out/pytest_frontend.log 2>&1
tail -30 gpurun_out/pytest_frontend.log
timeout 300 python tools/bench_configs.py --config 6 --cpu > gpurun_out/config6.jsonl 2> gpurun_out/config6.err
cat gpurun_out/config6.jsonl; tail -3 gpurun_out/config6.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mel_spectrogram_kernel -s 50 -c 1 \
    -o gpurun_out/mel_frontend python tools/bench_configs.py --config 6 > gpurun_out/ncu_mel.log 2>&1
ncu -i gpurun_out/mel_frontend.ncu-rep --page raw --csv > gpurun_out/mel_frontend_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
