#!/bin/bash
# Round 2, GPU job A: full GPU parity suite (incl. the full-size comparisons against the reference modules), the N=1 bench
# with both arms, the TF32 peak measurement, the launch list of one step and ncu --set full of the HBM-bound kernels
# north_star names (SCLN / LayerNorm, length-regulator gather, attention softmax).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
grep "\[parity\]" gpurun_out/pytest_gpu.log > gpurun_out/parity_fullsize.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 1500 gpurun_out/bench_reference.json
timeout 200 python tools/measure_tf32_peak.py > gpurun_out/tf32_peak.json 2> gpurun_out/tf32_peak.err; cat gpurun_out/tf32_peak.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/prof_step.py > gpurun_out/launches.log 2>&1
for k in layer_norm_kernel length_regulate_gather_kernel attn_softmax_warp_kernel; do
  timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:$k -c 3 \
      -o gpurun_out/ncu_$k python tools/prof_step.py > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/ncu_$k.ncu-rep --page raw --csv > gpurun_out/ncu_${k}_raw.csv 2>/dev/null
  rm -f gpurun_out/ncu_$k.ncu-rep
done
ls -la gpurun_out | tail -20
