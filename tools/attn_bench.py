"""Stand-alone timing of the fused attention kernel (C ABI zvx_attention) at the decoder shape of configs[1]
(B = 32, L = 821, 2 heads x 264), CUDA events, L2 flushed between launches.

    python tools/attn_bench.py [--B 32] [--L 821] [--heads 2] [--dk 264] [--iters 20]
Debug builds (ZVX_BUILD_DEBUG=1) with ZVX_ATTN_DBG=1 print the wait-cycle counters of CTA 0's roles after every launch.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from zerovox_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--L", type=int, default=821)
    ap.add_argument("--heads", type=int, default=2)
    ap.add_argument("--dk", type=int, default=264)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--min-len", type=float, default=0.85, help="utterance lengths ~ U{min_len * L .. L} (configs[1]: 700..821)")
    ap.add_argument("--plan", type=int, default=0, help="1 = device-built tile list, 2 = + masked query rows skipped")
    ap.add_argument("--variant", type=int, default=0, help="0 = library choice, 1 = single-CTA kernel, 2 = CTA-pair kernel")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    H, Lp = a.heads * a.dk, (a.L + 3) // 4 * 4
    g = torch.Generator(device=dev).manual_seed(0)
    qk = torch.randn((a.B * a.L, 2 * H), device=dev, generator=g)
    vt = torch.randn((a.B, H, Lp), device=dev, generator=g)
    out = torch.empty((a.B * a.L, H), device=dev)
    lens = torch.randint(int(a.L * a.min_len), a.L + 1, (a.B,), device=dev, generator=g)
    lens[0] = a.L
    mask = (torch.arange(a.L, device=dev)[None, :] >= lens[:, None]).to(torch.uint8).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream

    ws_bytes = lib.zvx_attention_workspace_bytes(a.B, a.L, a.heads) if a.plan else 0
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)

    def launch():
        rc = lib.zvx_attention_ex(C.c_void_p(qk.data_ptr()), C.c_void_p(vt.data_ptr()), Lp, C.c_void_p(mask.data_ptr()), a.B, a.L,
                                  a.heads, a.dk, C.c_float(a.dk ** 0.5), C.c_void_p(out.data_ptr()), a.variant,
                                  1 if a.plan == 2 else 0, C.c_void_p(ws.data_ptr()) if a.plan else None, ws_bytes, C.c_void_p(st))
        assert rc == 0, lib.zvx_attention_last_error().decode()

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    med = ms[len(ms) // 2]
    flops = 4.0 * a.B * a.heads * a.L * a.L * a.dk
    print(json.dumps({"kernel": "attn_fused", "variant": a.variant, "plan": a.plan, "valid_frac": float(lens.float().mean() / a.L), "B": a.B, "L": a.L, "heads": a.heads, "dk": a.dk, "us_median": med * 1e3,
                      "us_min": ms[0] * 1e3, "tflops": flops / (med * 1e-3) / 1e12}))


if __name__ == "__main__":
    main()
