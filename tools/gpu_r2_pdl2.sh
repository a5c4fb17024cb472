#!/bin/bash
# Round 2, GPU job: programmatic dependent launch with the LATE trigger in gemm_tc (start of a CTA's last tile) — A/B bench.
set -x
mkdir -p gpurun_out
for pdl in 1 0 1 0; do
timeout 300 python bench.py --steps 20 --warmup 5 --pdl $pdl --no-cpu-baseline --no-e2e > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err; python - $pdl <<'PY'
import json,sys
d=json.loads([x for x in open(f'gpurun_out/bench_pdl{sys.argv[1]}.json') if x.startswith('{')][-1])
print("pdl", sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"])
PY
done
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py -m gpu -q -x 2>&1 | tail -2
