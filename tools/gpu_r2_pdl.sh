#!/bin/bash
# Round 2, GPU job: programmatic dependent launch of the hot kernels — quick parity, A/B bench (pdl 1 / 0), then the whole suite.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3
for pdl in 1 0 1 0; do
timeout 300 python bench.py --steps 20 --warmup 5 --pdl $pdl --no-cpu-baseline > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err; python - $pdl <<'PY'
import json,sys
d=json.loads([x for x in open(f'gpurun_out/bench_pdl{sys.argv[1]}.json') if x.startswith('{')][-1])
print("pdl", sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"])
PY
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_pdl.log 2>&1; tail -3 gpurun_out/pytest_pdl.log
