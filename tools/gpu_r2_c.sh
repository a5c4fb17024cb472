#!/bin/bash
# Round 2, GPU job C: phase trace of the vocoder resblock kernels (debug build compiled on the box: -DZVX_DEBUG) + MMA ubench.
set -x
mkdir -p gpurun_out
timeout 120 ./build/ubench_mma > gpurun_out/ubench_mma_v2.txt 2>&1; cat gpurun_out/ubench_mma_v2.txt
timeout 120 ./build/ubench_mma --swz > gpurun_out/ubench_swz.txt 2>&1; cat gpurun_out/ubench_swz.txt
ZVX_BUILD_DEBUG=1 python __graft_entry__.py > gpurun_out/build_debug.log 2>&1; tail -2 gpurun_out/build_debug.log
ZVX_VOC_DBG=1 timeout 300 python tools/prof_step.py --warmup 1 2> gpurun_out/voc_dbg.txt > /dev/null
grep "voc dbg" gpurun_out/voc_dbg.txt | tail -9
