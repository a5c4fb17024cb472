#!/bin/bash
# Round 2: ncu --set full of the HBM-bound kernels north_star names, at their DECODER-sized launches of a configs[1] step
# (layer_norm_kernel: launches 15..17 are the decoder's SCLN; attn_softmax_warp_kernel<8>: the decoder's; the gather: its one launch).
set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:layer_norm_kernel --launch-skip 14 -c 3 \
    -o gpurun_out/ncu_ln python tools/prof_step.py > gpurun_out/ncu_ln.log 2>&1
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:attn_softmax_warp_kernel --launch-skip 4 -c 3 \
    -o gpurun_out/ncu_sm python tools/prof_step.py > gpurun_out/ncu_sm.log 2>&1
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:length_regulate_gather_kernel -c 1 \
    -o gpurun_out/ncu_lr python tools/prof_step.py > gpurun_out/ncu_lr.log 2>&1
timeout 300 ncu --set full --clock-control none --profile-from-start off -k "regex:conv_post_cl_kernel|se_scale_add_relu_kernel|hw_sum_partial_kernel" -c 5 \
    -o gpurun_out/ncu_misc python tools/prof_step.py > gpurun_out/ncu_misc.log 2>&1
for k in ln sm lr misc; do ncu -i gpurun_out/ncu_$k.ncu-rep --page raw --csv > gpurun_out/ncu_hbm_${k}_raw.csv 2>/dev/null; rm -f gpurun_out/ncu_$k.ncu-rep; done
python tools/ncu_summary.py gpurun_out/ncu_hbm_*_raw.csv > gpurun_out/ncu_hbm_summary.json; cat gpurun_out/ncu_hbm_summary.json | head -80
# the other BASELINE.json workloads through bench.py (one JSON line each) and a fresh full capture of every kernel class of a step
timeout 900 python bench.py --workload config3 --steps 5 > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; tail -c 600 gpurun_out/bench_config3.json
timeout 900 python bench.py --workload config5 --steps 5 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; tail -c 900 gpurun_out/bench_config5.json
timeout 900 ncu --set full --clock-control none --profile-from-start off -k "regex:gemm_tc_kernel|voc_pair_kernel" -c 160 \
    -o gpurun_out/ncu_step_tc python tools/prof_step.py > gpurun_out/ncu_step_tc.log 2>&1
ncu -i gpurun_out/ncu_step_tc.ncu-rep --page raw --csv > gpurun_out/ncu_step_tc_raw.csv 2>/dev/null; rm -f gpurun_out/ncu_step_tc.ncu-rep
python tools/ncu_summary.py gpurun_out/ncu_step_tc_raw.csv > gpurun_out/ncu_step_tc_summary.json; head -c 1500 gpurun_out/ncu_step_tc_summary.json
