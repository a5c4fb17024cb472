#!/usr/bin/env python
"""TF32 tensor peak on this B200, measured the way MEASURED_PEAKS.json measures bf16 (VERDICT r1 item 5 / BASELINE.md):

  * cuBLAS TF32: torch.matmul of two 8192^2 fp32 matrices with allow_tf32 (2*N^3 FLOP) — best of 10 (burst) and back to
    back for 4 s (sustained);
  * the tensor pipe's own ceiling: tcgen05.mma.kind::tf32 (M=128, N=256, K=8, 128B-swizzled shared-memory operands, no
    loads) issued back to back on every SM (tools/ubench_mma.cu --peak; build/ubench_mma is compiled in the build container).

    python tools/measure_tf32_peak.py > profiles/r02_tf32_peak.json
"""
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cublas_tf32():
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    flop = 2.0 * n ** 3
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0, k = time.perf_counter(), 0
    while time.perf_counter() - t0 < 4.0:
        for _ in range(20):
            a @ b
        k += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return flop / (best * 1e-3) / 1e12, flop * k / (e0.elapsed_time(e1) * 1e-3) / 1e12


def main():
    out = {"gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS TF32): best of 10 (burst), back to back for 4 s "
                  "(sustained); tools/ubench_mma.cu --peak: tcgen05.mma.kind::tf32 N=256 issue rate on all SMs"}
    burst, sust = cublas_tf32()
    out["cublas_tf32_tflops_burst"], out["cublas_tf32_tflops_sustained"] = round(burst, 1), round(sust, 1)
    exe = os.path.join(ROOT, "build", "ubench_mma")
    if os.path.exists(exe):
        try:
            r = subprocess.run([exe, "--peak"], capture_output=True, text=True, timeout=120)
            out.update(json.loads(r.stdout.strip().splitlines()[-1]))
        except Exception as e:  # noqa: BLE001
            out["ubench_error"] = repr(e)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
