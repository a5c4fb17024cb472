#!/bin/bash
# GPU job: parity tests, bench, ncu launch list, ncu --set full of the dominant kernels (reports kept < 64 MiB).
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json
if [ "$1" == "quick" ]; then exit 0; fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/prof_step.py > gpurun_out/launches.log 2>&1
# all 27 fused-pair launches: raw metrics only (the report itself is too large to bring back)
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:voc_pair_kernel \
    -o gpurun_out/voc_all python tools/prof_step.py > gpurun_out/ncu_voc.log 2>&1
ncu -i gpurun_out/voc_all.ncu-rep --page raw --csv > gpurun_out/voc_all_raw.csv 2>/dev/null
rm -f gpurun_out/voc_all.ncu-rep
# one launch per channel count with source (k = 11, d = 1)
for s in 6 15 24; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:voc_pair_kernel \
      -s $s -c 1 -o gpurun_out/voc_pair_s$s python tools/prof_step.py > gpurun_out/ncu_voc_s$s.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel \
    -s 65 -c 9 -o gpurun_out/gemm_tc_dec python tools/prof_step.py > gpurun_out/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tc_kernel \
    -s 31 -c 7 -o gpurun_out/gemm_tc_enc python tools/prof_step.py > gpurun_out/ncu_gemm_enc.log 2>&1
ncu -i gpurun_out/gemm_tc_enc.ncu-rep --page raw --csv > gpurun_out/gemm_tc_enc_raw.csv 2>/dev/null
ncu -i gpurun_out/gemm_tc_dec.ncu-rep --page raw --csv > gpurun_out/gemm_tc_dec_raw.csv 2>/dev/null
ls -la gpurun_out
du -sh gpurun_out
