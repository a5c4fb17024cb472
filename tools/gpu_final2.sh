#!/bin/bash
# Round-end GPU job of the final build: all parity tests, both bench arms, smoke, launch list of one step, memcheck of the attention
# kernels and of the speaker net / vocoder tail (squeeze+excite ticket kernel, register-weight conv_post), ncu --set full of the
# tensor-core kernels of a step + of the HBM-bound kernels north_star names, config3 / config5 bench lines.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 700 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python tools/bench_configs.py --config 2 > gpurun_out/stage_split.jsonl 2>&1; tail -1 gpurun_out/stage_split.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/prof_step.py > gpurun_out/launches.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "53 or 257 or 300-4" \
    > gpurun_out/sanitizer_attention.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/sanitizer_attention.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(speaker_embedding and (tiny or 48)) or (vocoder_variants and v2 and (9 or 70))" \
    > gpurun_out/sanitizer_spk_voc.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/sanitizer_spk_voc.log
timeout 900 ncu --set full --clock-control none --profile-from-start off -k "regex:gemm_tc_kernel|voc_pair_kernel|attn_" -c 170 \
    -o gpurun_out/ncu_step_tc python tools/prof_step.py > gpurun_out/ncu_step_tc.log 2>&1
ncu -i gpurun_out/ncu_step_tc.ncu-rep --page raw --csv > gpurun_out/ncu_step_tc_raw.csv 2>/dev/null; rm -f gpurun_out/ncu_step_tc.ncu-rep
python tools/ncu_summary.py gpurun_out/ncu_step_tc_raw.csv > gpurun_out/ncu_step_tc_summary.json; head -c 600 gpurun_out/ncu_step_tc_summary.json
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:layer_norm_kernel --launch-skip 14 -c 3 \
    -o gpurun_out/ncu_ln python tools/prof_step.py > gpurun_out/ncu_ln.log 2>&1
timeout 300 ncu --set full --clock-control none --profile-from-start off -k "regex:conv_post_cl|se_scale_add_relu_kernel|se_squeeze_excite_kernel|length_regulate_gather_kernel" -c 7 \
    -o gpurun_out/ncu_misc python tools/prof_step.py > gpurun_out/ncu_misc.log 2>&1
for k in ln misc; do ncu -i gpurun_out/ncu_$k.ncu-rep --page raw --csv > gpurun_out/ncu_hbm2_${k}_raw.csv 2>/dev/null; rm -f gpurun_out/ncu_$k.ncu-rep; done
python tools/ncu_summary.py gpurun_out/ncu_hbm2_*_raw.csv > gpurun_out/ncu_hbm2_summary.json; head -c 1200 gpurun_out/ncu_hbm2_summary.json
timeout 600 python bench.py --workload config3 --steps 5 > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; tail -c 400 gpurun_out/bench_config3.json
timeout 600 python bench.py --workload config5 --steps 5 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; tail -c 600 gpurun_out/bench_config5.json
ls -la gpurun_out | tail -20
