"""GPU micro-benchmark of the tcgen05 GEMM kernel (zvx_debug_gemm) on the engine's hot shapes; CUDA-event timing."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zerovox_b200 import _lib
from zerovox_b200.engine import Engine, EngineConfig
eng = Engine(EngineConfig(), "cuda:0")
DEV = "cuda:0"

def bench(name, mode, M, N, K, taps=1, bias=True, res=False, L=0, Hh=0, Ww=0, ksize=1, pad=0, iters=20):
    A = torch.randn(M, K, device=DEV); W = torch.randn(taps, N, K, device=DEV) * 0.05
    b = torch.randn(N, device=DEV) if bias else None
    R = torch.randn(M, N, device=DEV) if res else None
    out = torch.empty(M, N, device=DEV)
    d = _lib.ZvxGemmDesc()
    d.A, d.W, d.C = A.data_ptr(), W.data_ptr(), out.data_ptr()
    d.bias = b.data_ptr() if bias else None; d.R = R.data_ptr() if res else None
    d.M, d.N, d.K, d.taps, d.mode, d.L, d.Hh, d.Ww, d.ksize, d.pad, d.dil = M, N, K, taps, mode, L, Hh, Ww, ksize, pad, 1
    d.lda, d.ldw, d.ldc = K, K, N
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = eng.lib.zvx_debug_gemm(eng._h, C.byref(d), 1, None)
        e.record()
        assert rc == 0, eng.lib.zvx_last_error(eng._h).decode()
        torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    fl = 2.0 * M * N * K * taps
    by = 4.0 * (M * K + N * K * taps + M * N * (2 if res else 1))
    print(f"{name:28s} M={M:6d} N={N:5d} K={K:5d} taps={taps:2d}  {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TF/s  {by/ms/1e6:7.1f} GB/s")

B, L = 32, 821
bench("dec QK proj", 0, B * L, 1056, 528)
bench("dec fc", 0, B * L, 528, 528)
bench("dec fc + residual", 0, B * L, 528, 528, res=True)
bench("dec w1 conv k9", 1, B * L, 1024, 528, taps=9, L=L, pad=4)
bench("dec w2", 0, B * L, 528, 1024)
bench("dec mel_linear", 0, B * L, 80, 528)
bench("big square", 0, 8192, 8192, 8192, bias=False, iters=5)
bench("spk conv 32ch 80x440", 2, B * 80 * 440, 32, 32, taps=9, Hh=80, Ww=440, ksize=3, pad=1)
bench("spk conv 64ch 40x220", 2, B * 40 * 220, 64, 64, taps=9, Hh=40, Ww=220, ksize=3, pad=1)
bench("spk conv 128ch 20x110", 2, B * 20 * 110, 128, 128, taps=9, Hh=20, Ww=110, ksize=3, pad=1)
bench("spk conv 256ch 10x55", 2, B * 10 * 55, 256, 256, taps=9, Hh=10, Ww=55, ksize=3, pad=1)
bench("voc s1 conv 64ch k11", 1, B * L * 8, 64, 64, taps=11, L=L * 8, pad=5)
