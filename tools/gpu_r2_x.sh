#!/bin/bash
# Round 2, GPU job X: fused SE squeeze + excitation — speaker-net parity, determinism across engines, stage split; large-B plan test.
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "speaker or forward or smoke" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "1100" 2>&1 | tail -3
timeout 300 python tools/bench_configs.py --config 2 2>&1 | tail -1
