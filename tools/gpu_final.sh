#!/bin/bash
# Round-end GPU job (short form of gpu_job.sh): all parity tests, both bench arms, smoke, launch list of one step,
# memcheck of the front-end kernels, ncu --set full of the first gemm_tc launches of a step (speaker net, two-M-tile path).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/prof_step.py > gpurun_out/launches.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_frontend.py -m gpu -q -x \
    > gpurun_out/sanitizer_frontend.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/sanitizer_frontend.log
timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -c 16 \
    -o gpurun_out/gemm_tc_spk python tools/prof_step.py > gpurun_out/ncu_gemm_spk.log 2>&1
ncu -i gpurun_out/gemm_tc_spk.ncu-rep --page raw --csv > gpurun_out/gemm_tc_spk_raw.csv 2>/dev/null
rm -f gpurun_out/gemm_tc_spk.ncu-rep
ls -la gpurun_out | tail -12
