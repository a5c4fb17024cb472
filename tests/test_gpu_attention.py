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


def run(q, k, v, mask, temperature, variant=0, mode="static"):
    """mode: "static" = every tile of the padded batch; "plan" = device-built tile list (masked key blocks skipped);
    "plan_skip" = query rows past the last unmasked position are not computed either (left as NaN here)"""
    B, L, nh, dk = q.shape
    H = nh * dk
    Lp = (L + 3) // 4 * 4
    qk = torch.cat([q.reshape(B * L, H), k.reshape(B * L, H)], dim=1).to(DEV).contiguous()
    vt = torch.zeros((B, H, Lp), device=DEV)
    vt[:, :, :L] = v.reshape(B, L, H).transpose(1, 2).to(DEV)
    out = torch.full((B * L, H), float("nan"), device=DEV)
    m = mask.to(torch.uint8).to(DEV).contiguous() if mask is not None else None
    lib = _lib.load()
    ws, ws_bytes = None, 0
    if mode != "static":
        ws_bytes = lib.zvx_attention_workspace_bytes(B, L, nh)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
    rc = lib.zvx_attention_ex(C.c_void_p(qk.data_ptr()), C.c_void_p(vt.data_ptr()), Lp, C.c_void_p(m.data_ptr()) if m is not None else None,
                              B, L, nh, dk, C.c_float(temperature), C.c_void_p(out.data_ptr()), variant,
                              1 if mode == "plan_skip" else 0, C.c_void_p(ws.data_ptr()) if ws is not None else None, ws_bytes,
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
    (37, 300, 2, 264, True, 1.0),     # pair kernel: 148 pair tiles on 74 clusters (Q replaced between tiles), odd q-tile count
    (1, 257, 2, 128, True, 1.0),      # pair kernel: second pair holds one row
    (9, 640, 2, 264, True, 1.0),      # ragged: utterances with 2..5 q tiles and 2..6 key blocks in one tile list
    (1100, 150, 1, 32, True, 1.0),    # more utterances than the plan kernel has threads (its scan runs in rounds of 1024)
]


# variant 1 = single-CTA kernel, 2 = CTA-pair kernel (tcgen05.mma.cta_group::2, Q resident), 0 = the library's choice
VARIANTS = [(1, "static"), (1, "plan"), (1, "plan_skip"), (2, "static"), (2, "plan"), (2, "plan_skip"), (0, "static"), (0, "plan_skip")]
VARIANT_IDS = ["single", "single-plan", "single-skip", "pair", "pair-plan", "pair-skip", "auto", "auto-skip"]


@pytest.mark.parametrize("variant,mode", VARIANTS, ids=VARIANT_IDS)
@pytest.mark.parametrize("B,L,nh,dk,ragged,gain", CASES)
def test_attention_matches_float64(B, L, nh, dk, ragged, gain, variant, mode):
    g = torch.Generator().manual_seed(B * 1000 + L + dk)
    q = torch.randn((B, L, nh, dk), generator=g) * gain
    k = torch.randn((B, L, nh, dk), generator=g)
    v = torch.randn((B, L, nh, dk), generator=g)
    mask = None
    valid = torch.ones((B, L, 1), dtype=torch.bool)      # rows the kernel has to produce
    if ragged:
        lens = torch.randint(max(1, L // 3), L + 1, (B,), generator=g)
        lens[0] = L
        mask = torch.arange(L)[None, :] >= lens[:, None]
        if mode == "plan_skip":
            valid = ~mask[:, :, None]
    T = float(dk) ** 0.5

    def clean(x):
        return torch.where(valid, x, torch.zeros(()))

    out = clean(run(q, k, v, mask, T, variant, mode))
    ref = clean(reference(q.double(), k.double(), v.double(), mask, T).float())
    scale = float(ref.abs().max())
    err = float((out - ref).abs().max())
    # peaked softmax: the score error (2^-11 * |q||k| / T) is amplified by the gain
    bar = 2e-3 * scale * max(1.0, gain / 4)
    assert torch.isfinite(out).all()
    assert err <= bar, f"max err {err:.3e} > {bar:.3e} (scale {scale:.3f})"
    # operands that are already TF32 numbers: only P's rounding and fp32 accumulation differ
    qt, kt, vt = to_tf32(q), to_tf32(k), to_tf32(v)
    out2 = clean(run(qt, kt, vt, mask, T, variant, mode))
    ref2 = clean(reference(qt.double(), kt.double(), vt.double(), mask, T).float())
    err2 = float((out2 - ref2).abs().max())
    assert err2 <= 6e-4 * float(ref2.abs().max()), f"TF32-exact operands: max err {err2:.3e}"


def test_attention_rejects_bad_shapes():
    lib = _lib.load()
    x = torch.zeros(64, device=DEV)
    rc = lib.zvx_attention(C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), 4, None, 1, 4, 1, 6, C.c_float(1.0),
                           C.c_void_p(x.data_ptr()), None)
    assert rc != 0 and b"unsupported" in lib.zvx_attention_last_error()
    # a workspace that is too small, and skipping without one
    qk = torch.zeros((128, 128), device=DEV)
    rc = lib.zvx_attention_ex(C.c_void_p(qk.data_ptr()), C.c_void_p(qk.data_ptr()), 128, None, 1, 128, 1, 64, C.c_float(8.0),
                              C.c_void_p(qk.data_ptr()), 1, 0, C.c_void_p(x.data_ptr()), 16, None)
    assert rc != 0 and b"workspace" in lib.zvx_attention_last_error()
    rc = lib.zvx_attention_ex(C.c_void_p(qk.data_ptr()), C.c_void_p(qk.data_ptr()), 128, None, 1, 128, 1, 64, C.c_float(8.0),
                              C.c_void_p(qk.data_ptr()), 1, 1, None, 0, None)
    assert rc != 0 and b"workspace" in lib.zvx_attention_last_error()
