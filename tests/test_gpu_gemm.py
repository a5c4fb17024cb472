"""Kernel-level numerics of the two contraction kernels (fp32 FMA and tcgen05 TF32) through the C-ABI test hook
zvx_debug_gemm, against a plain PyTorch fp32 reference of the same op (torch CPU float64 accumulate -> fp32).

Tolerances: fp32 FMA kernel |err| <= 1e-5 * sqrt(K_total) * max|ref|-scale; TF32 kernel: operands carry 10 mantissa
bits (rel 2^-10 truncation), so |err| <= 2e-3 * ||a_row|| * ||w_col|| bound, checked as 4e-3 * max|ref| for the
N(0,1) operands used here.  The 3xTF32 split mode (use_tc=2: hi*hi + hi*lo + lo*hi, fp32 accumulate) is held to the
fp32 FMA bar.
"""
import ctypes as C

import pytest
import torch

from zerovox_b200 import _lib
from zerovox_b200.engine import Engine, EngineConfig

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def eng():
    return Engine(EngineConfig(), DEV)


def run(eng, mode, A, W, bias=None, scale=None, shift=None, R=None, relu_first=0, relu_last=0, use_tc=1, L=0, Hh=0,
        Ww=0, ksize=1, pad=0, dil=1, stride=1, M_out=None):
    """A [M, K]; W [taps, N, K] -> C [M, N] on the device."""
    M, K = A.shape
    if M_out is not None:
        M = M_out
    taps, N, _ = W.shape
    Cout = torch.full((M, N), float("nan"), device=DEV)
    keep = [t.to(DEV).contiguous() if t is not None else None for t in (A, W, bias, scale, shift, R)]
    d = _lib.ZvxGemmDesc()
    for name, t in zip(("A", "W", "bias", "scale", "shift", "R"), keep):
        setattr(d, name, t.data_ptr() if t is not None else None)
    d.C = Cout.data_ptr()
    d.M, d.N, d.K, d.taps, d.mode, d.L, d.Hh, d.Ww, d.ksize, d.pad, d.dil = M, N, K, taps, mode, L, Hh, Ww, ksize, pad, dil
    d.relu_first, d.relu_last, d.lda, d.ldw, d.ldc = relu_first, relu_last, K, K, N
    d.stride = stride
    rc = eng.lib.zvx_debug_gemm(eng._h, C.byref(d), use_tc, None)
    assert rc == 0, eng.lib.zvx_last_error(eng._h).decode()
    torch.cuda.synchronize()
    return Cout.cpu()


def epilogue(y, bias, scale, shift, R, relu_first, relu_last):
    if bias is not None:
        y = y + bias
    if relu_first:
        y = y.relu()
    if scale is not None:
        y = y * scale + shift
    if R is not None:
        y = y + R
    if relu_last:
        y = y.relu()
    return y


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def check(got, ref, use_tc, ktot):
    assert torch.isfinite(got).all(), "kernel left NaN / unwritten outputs"
    err = (got.double() - ref.double()).abs().max().item()
    mag = ref.abs().max().item()
    # use_tc: 0 fp32 FMA; 1 TF32 tensor cores; 2 3xTF32 split on the tensor cores (fp32-grade products: same bar as FMA)
    tol = (4e-3 if use_tc == 1 else 2e-5) * max(mag, ktot ** 0.5)
    print(f"  max|diff|={err:.3e} max|ref|={mag:.3e} tol={tol:.3e}")
    assert err <= tol


@pytest.mark.parametrize("use_tc", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(128, 16, 32), (128, 176, 528), (300, 528, 528), (1000, 1056, 528), (257, 80, 528),
                                   (4096, 1024, 264), (64, 24, 40), (777, 130, 100)])
def test_plain_gemm(eng, M, N, K, use_tc):
    A, W = rnd(M, K, seed=1), rnd(1, N, K, seed=2)
    bias, R = rnd(N, seed=3), rnd(M, N, seed=4)
    got = run(eng, 0, A, W, bias=bias, R=R, relu_last=1, use_tc=use_tc)
    ref = epilogue((A.double() @ W[0].double().T).float(), bias, None, None, R, 0, 1)
    check(got, ref, use_tc, K)


@pytest.mark.parametrize("use_tc", [0, 1, 2])
@pytest.mark.parametrize("B,L,Cin,Cout,k,dil", [(2, 128, 64, 32, 3, 1), (3, 200, 528, 1024, 9, 1), (1, 821, 528, 256, 9, 1),
                                                (5, 77, 256, 256, 3, 1), (4, 100, 32, 48, 5, 2)])
def test_conv1d(eng, B, L, Cin, Cout, k, dil, use_tc):
    x, w = rnd(B, L, Cin, seed=5), rnd(Cout, Cin, k, seed=6) / (Cin * k) ** 0.5
    bias = rnd(Cout, seed=7)
    pad = dil * (k - 1) // 2
    Wt = w.permute(2, 0, 1).contiguous()  # tap-major [k][Cout][Cin]
    got = run(eng, 1, x.reshape(B * L, Cin), Wt, bias=bias, relu_first=1, use_tc=use_tc, L=L, pad=pad, dil=dil)
    ref = torch.nn.functional.conv1d(x.double().transpose(1, 2), w.double(), bias.double(), padding=pad, dilation=dil)
    ref = ref.relu().transpose(1, 2).reshape(B * L, Cout).float()
    check(got, ref, use_tc, 1.0)


@pytest.mark.parametrize("B,L,Cin,Cout,k,dil", [(12, 8192, 64, 64, 7, 3), (10, 8200, 80, 128, 7, 1), (16, 6568, 64, 64, 11, 5),
                                                (9, 12001, 32, 16, 5, 1)])
def test_conv1d_tap_reuse_two_m_tiles(eng, B, L, Cin, Cout, k, dil):
    """Filter-row tap reuse with two M tiles per CTA tile (256 + halo positions behind one weight stage): shapes large
    enough to select it — HiFi-GAN's 64-channel k = 7 / 11 dilated convs, conv_pre's 80 -> 128, a ragged row length."""
    x, w = rnd(B, L, Cin, seed=30), rnd(Cout, Cin, k, seed=31) / (Cin * k) ** 0.5
    bias = rnd(Cout, seed=32)
    pad = dil * (k - 1) // 2
    Wt = w.permute(2, 0, 1).contiguous()
    got = run(eng, 1, x.reshape(B * L, Cin), Wt, bias=bias, relu_first=1, use_tc=1, L=L, pad=pad, dil=dil)
    ref = torch.nn.functional.conv1d(x.double().transpose(1, 2), w.double(), bias.double(), padding=pad, dilation=dil)
    check(got, ref.relu().transpose(1, 2).reshape(B * L, Cout).float(), 1, 1.0)


@pytest.mark.parametrize("use_tc", [0, 1, 2])
@pytest.mark.parametrize("B,Hh,Ww,Cin,Cout", [(2, 80, 48, 32, 32), (1, 40, 33, 64, 64), (3, 10, 7, 256, 256),
                                              (2, 20, 55, 128, 128), (3, 80, 440, 32, 32), (5, 40, 220, 64, 64)])
def test_conv2d_3x3(eng, B, Hh, Ww, Cin, Cout, use_tc):
    x, w = rnd(B, Hh, Ww, Cin, seed=8), rnd(Cout, Cin, 3, 3, seed=9) / (Cin * 9) ** 0.5
    scale, shift = rnd(Cout, seed=10).abs() + 0.5, rnd(Cout, seed=11)
    Wt = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
    got = run(eng, 2, x.reshape(-1, Cin), Wt, scale=scale, shift=shift, relu_first=1, use_tc=use_tc, Hh=Hh, Ww=Ww,
              ksize=3, pad=1)
    ref = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double(), padding=1).relu()
    ref = (ref * scale.double()[None, :, None, None] + shift.double()[None, :, None, None]).permute(0, 2, 3, 1)
    check(got, ref.reshape(-1, Cout).float(), use_tc, 1.0)


@pytest.mark.parametrize("B,Hh,Ww,Cin,Cout,ksize", [(2, 80, 440, 16, 16, 5), (2, 64, 300, 32, 48, 3), (1, 96, 410, 40, 24, 5)])
def test_conv2d_tap_reuse_other_filters(eng, B, Hh, Ww, Cin, Cout, ksize):
    """The 2-D tap-reuse path of gemm_tc (one haloed activation tile per k-chunk) on shapes large enough to select it:
    5x5 filters (halo 4), channel counts that are not multiples of 32 (k-chunk tail), Cout that needs a column tail."""
    x, w = rnd(B, Hh, Ww, Cin, seed=20), rnd(Cout, Cin, ksize, ksize, seed=21) / (Cin * ksize * ksize) ** 0.5
    bias = rnd(Cout, seed=22)
    Wt = w.permute(2, 3, 0, 1).reshape(ksize * ksize, Cout, Cin).contiguous()
    got = run(eng, 2, x.reshape(-1, Cin), Wt, bias=bias, relu_last=1, use_tc=1, Hh=Hh, Ww=Ww, ksize=ksize, pad=ksize // 2)
    ref = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double(), bias.double(), padding=ksize // 2).relu()
    check(got, ref.permute(0, 2, 3, 1).reshape(-1, Cout).float(), 1, 1.0)


@pytest.mark.parametrize("use_tc", [0, 1])
@pytest.mark.parametrize("B,Hh,Ww,Cin,Cout,ksize", [(2, 80, 48, 32, 64, 3), (3, 40, 55, 64, 128, 3), (2, 20, 27, 128, 256, 3),
                                                    (2, 80, 48, 32, 64, 1), (1, 21, 33, 64, 128, 1)])
def test_conv2d_stride2(eng, B, Hh, Ww, Cin, Cout, ksize, use_tc):
    """Strided stage-entry convolutions of ResNetSE34V2 (3x3 pad 1 and the 1x1 downsample, ResNetSE34V2.py:81-92, 161-167):
    on the tensor-core path the stride is a TMA element stride."""
    pad = ksize // 2
    x, w = rnd(B, Hh, Ww, Cin, seed=12), rnd(Cout, Cin, ksize, ksize, seed=13) / (Cin * ksize * ksize) ** 0.5
    scale, shift = rnd(Cout, seed=14).abs() + 0.5, rnd(Cout, seed=15)
    Wt = w.permute(2, 3, 0, 1).reshape(ksize * ksize, Cout, Cin).contiguous()
    ref = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double(), padding=pad, stride=2)
    Ho, Wo = ref.shape[2], ref.shape[3]
    ref = (ref * scale.double()[None, :, None, None] + shift.double()[None, :, None, None]).permute(0, 2, 3, 1)
    got = run(eng, 2, x.reshape(-1, Cin), Wt, scale=scale, shift=shift, use_tc=use_tc, Hh=Hh, Ww=Ww, ksize=ksize, pad=pad,
              stride=2, M_out=B * Ho * Wo)
    check(got, ref.reshape(-1, Cout).float(), use_tc, 1.0)


@pytest.mark.parametrize("M,N,K", [(32, 528, 5120), (5, 1056, 528), (64, 100, 300), (1, 528, 5120), (33, 8, 256),
                                   (32, 12672, 528), (7, 4230, 300), (40, 8192, 256)])
def test_skinny_gemm(eng, M, N, K):
    """M <= 64 rows (speaker-net fc 5120 -> 528, SCLN affine stack): the fp32 FMA path switches to the skinny kernel."""
    A, W = rnd(M, K, seed=21), rnd(1, N, K, seed=22)
    bias, R = rnd(N, seed=23), rnd(M, N, seed=24)
    got = run(eng, 0, A, W, bias=bias, R=R, relu_last=1, use_tc=0)
    ref = epilogue((A.double() @ W[0].double().T).float(), bias, None, None, R, 0, 1)
    check(got, ref, 0, K)
