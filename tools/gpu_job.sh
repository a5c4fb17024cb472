#!/bin/bash
# Full GPU job: parity tests, bench (both arms), ncu launch list, ncu --set full of the dominant kernels.
# Reports are reduced on the box (raw-metric CSVs) so that gpurun_out/ stays below the 64 MiB it can bring back.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json
if [ "$1" == "quick" ]; then exit 0; fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/prof_step.py > gpurun_out/launches.log 2>&1
# every fused-resblock launch: raw metrics only
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:voc_poly_kernel \
    -o gpurun_out/voc_all python tools/prof_step.py > gpurun_out/ncu_voc.log 2>&1
ncu -i gpurun_out/voc_all.ncu-rep --page raw --csv > gpurun_out/voc_poly_raw.csv 2>/dev/null
rm -f gpurun_out/voc_all.ncu-rep
# one launch with source (C = 8, k = 11)
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:voc_poly_kernel \
    -s 8 -c 1 -o gpurun_out/voc_poly_c8k11 python tools/prof_step.py > gpurun_out/ncu_voc_src.log 2>&1
# every launch of the TMA GEMM family (speaker net, encoder 3xTF32, decoder, vocoder): raw metrics only
timeout 1500 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tc_kernel \
    -o gpurun_out/gemm_tc_all python tools/prof_step.py > gpurun_out/ncu_gemm.log 2>&1
ncu -i gpurun_out/gemm_tc_all.ncu-rep --page raw --csv > gpurun_out/gemm_tc_all_raw.csv 2>/dev/null
rm -f gpurun_out/gemm_tc_all.ncu-rep
# the FFN k = 9 conv of decoder layer 0 with source (largest single kernel)
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel \
    -s 76 -c 1 -o gpurun_out/gemm_tc_ffn_conv python tools/prof_step.py > gpurun_out/ncu_gemm_src.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
ls -la gpurun_out
du -sh gpurun_out
