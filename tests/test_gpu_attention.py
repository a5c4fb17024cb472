"""The fused attention kernel (csrc/attn_fused.cu, C ABI zvx_attention) against a float64 PyTorch statement of
fs2.py:101-163 (bmm, / temperature, masked_fill(-inf), softmax, bmm) on the same operands.

Tolerance: both products run in TF32 (operands rounded to 10 mantissa bits, fp32 accumulate), the softmax in fp32.  With
N(0,1) operands a score carries ~2^-11 * sqrt(d_k) absolute error before the division by temperature = sqrt(d_k), i.e. ~5e-4 in
the exponent, and P / V another 2^-11 relative each: |err| <= 2e-3 * max|ref| is the bar (measured: 3-6e-4); the same
operands pre-rounded to TF32 on the host must agree to fp32 accumulation noise (2e-5 * max|ref|).
"""
import ctypes as C

import pytest
import torch

from zerovox_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def to_tf32(x):
    """round-to-nearest (ties away, cvt.rna) to 10 mantissa bits"""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def reference(q, k, v, mask, temperature):
    """q, k, v: [B, L, nh, dk] float64; mask [B, L] bool (True = masked key) -> [B, L, nh*dk]"""
    B, L, nh, dk = q.shape
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) / temperature
    if mask is not None:
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v).reshape(B, L, nh * dk)


def run(q, k, v, mask, temperature):
    B, L, nh, dk = q.shape
    H = nh * dk
    Lp = (L + 3) // 4 * 4
    qk = torch.cat([q.reshape(B * L, H), k.reshape(B * L, H)], dim=1).to(DEV).contiguous()
    vt = torch.zeros((B, H, Lp), device=DEV)
    vt[:, :, :L] = v.reshape(B, L, H).transpose(1, 2).to(DEV)
    out = torch.full((B * L, H), float("nan"), device=DEV)
    m = mask.to(torch.uint8).to(DEV).contiguous() if mask is not None else None
    lib = _lib.load()
    rc = lib.zvx_attention(C.c_void_p(qk.data_ptr()), C.c_void_p(vt.data_ptr()), Lp, C.c_void_p(m.data_ptr()) if m is not None else None,
                           B, L, nh, dk, C.c_float(temperature), C.c_void_p(out.data_ptr()),
                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.zvx_attention_last_error().decode()
    torch.cuda.synchronize()
    return out.view(B, L, H).cpu()


CASES = [
    # B, L, nh, dk, ragged mask, score gain
    (2, 821, 2, 264, True, 1.0),      # configs[1] decoder shape
    (3, 128, 2, 264, False, 1.0),     # exactly one key block
    (2, 53, 2, 264, True, 1.0),       # less than one block
    (1, 1000, 1, 64, True, 1.0),      # single V tile (NV <= 256)
    (2, 300, 4, 256, False, 1.0),
    (1, 700, 2, 264, True, 40.0),     # peaked scores: the lazy reference maximum moves, O is rescaled in TMEM
    (40, 200, 2, 264, True, 1.0),     # more tiles than SMs: the persistent loop, TMEM / barrier phases across tiles
]


@pytest.mark.parametrize("B,L,nh,dk,ragged,gain", CASES)
def test_attention_matches_float64(B, L, nh, dk, ragged, gain):
    g = torch.Generator().manual_seed(B * 1000 + L + dk)
    q = torch.randn((B, L, nh, dk), generator=g) * gain
    k = torch.randn((B, L, nh, dk), generator=g)
    v = torch.randn((B, L, nh, dk), generator=g)
    mask = None
    if ragged:
        lens = torch.randint(max(1, L // 3), L + 1, (B,), generator=g)
        lens[0] = L
        mask = torch.arange(L)[None, :] >= lens[:, None]
    T = float(dk) ** 0.5
    out = run(q, k, v, mask, T)
    ref = reference(q.double(), k.double(), v.double(), mask, T).float()
    scale = float(ref.abs().max())
    err = float((out - ref).abs().max())
    # peaked softmax: the score error (2^-11 * |q||k| / T) is amplified by the gain
    bar = 2e-3 * scale * max(1.0, gain / 4)
    assert torch.isfinite(out).all()
    assert err <= bar, f"max err {err:.3e} > {bar:.3e} (scale {scale:.3f})"
    # operands that are already TF32 numbers: only P's rounding and fp32 accumulation differ
    qt, kt, vt = to_tf32(q), to_tf32(k), to_tf32(v)
    out2 = run(qt, kt, vt, mask, T)
    ref2 = reference(qt.double(), kt.double(), vt.double(), mask, T).float()
    err2 = float((out2 - ref2).abs().max())
    assert err2 <= 6e-4 * float(ref2.abs().max()), f"TF32-exact operands: max err {err2:.3e}"


def test_attention_rejects_bad_shapes():
    lib = _lib.load()
    x = torch.zeros(64, device=DEV)
    rc = lib.zvx_attention(C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), 4, None, 1, 4, 1, 6, C.c_float(1.0),
                           C.c_void_p(x.data_ptr()), None)
    assert rc != 0 and b"unsupported" in lib.zvx_attention_last_error()
