#!/bin/bash
# Round 2, GPU job S (debug build): fused attention — wait counters / timeline of CTA 0 for a kernel variant ($1, default 2).
V=${1:-2}
ZVX_ATTN_DBG=1 timeout 200 python tools/attn_bench.py --iters 2 --variant $V 2>&1 | tail -11
timeout 200 python tools/attn_bench.py --variant $V 2>&1 | tail -1
