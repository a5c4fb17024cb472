#!/bin/bash
# Round 2, GPU job W: bench line of the current build (N = 1) with the pipelined host delivery variants; full GPU suite.
set -x
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2500 gpurun_out/bench_n1.json | head -c 1200; echo; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_n1.json') if x.startswith('{')][-1]
d=json.loads(l)
print(d["value"], d["ms_per_step"], json.dumps(d["e2e"])[:900])
print(json.dumps(d["roofline"])[:600])
PY
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
