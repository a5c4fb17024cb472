#!/bin/bash
# Round 2, GPU job Y: zvx_spkemb_encode (speaker net on the engine's side stream next to the encoder) — equivalence test, forward
# parity, stage split (separate calls) against the step (one call), bench line.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_patch.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/bench_configs.py --config 2 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; python - <<'PY'
import json
d=json.loads([x for x in open('gpurun_out/bench_y.json') if x.startswith('{')][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"])
PY
