#!/bin/bash
# Round 2, GPU job Z: squeeze+excite in one launch, layer_norm sized to the row, conv_post with register weights, duration predictor
# on the side stream — the whole GPU suite, the stage split and the bench line.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_z.log 2>&1; tail -3 gpurun_out/pytest_z.log
timeout 300 python tools/bench_configs.py --config 2 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err; python - <<'PY'
import json
d=json.loads([x for x in open('gpurun_out/bench_z.json') if x.startswith('{')][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"])
PY
