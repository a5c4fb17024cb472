"""GPU side of the multi-GPU boundary (zerovox_b200/parallel.py) that the gloo tests cannot reach: the ragged pack / unpack
kernels of csrc/ragged.cu through the C ABI, and `sharded_forward` on one GPU (world size 1: no collective, but the same
packing, `RaggedBatch`, grouped vocoding and host delivery code) against a plain `ZeroVox.forward` call."""
import numpy as np
import pytest
import torch

from oracle import zerovox_oracle as zo
from zerovox_b200.parallel import _ragged_copy, sharded_forward
from zerovox_b200.testing import build_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("rows,unit,max_units", [(1, 256, 37), (80, 1, 131), (1, 4, 9), (3, 6, 50)])
def test_ragged_pack_unpack_kernels(rows, unit, max_units):
    """zvx_ragged_pack / zvx_ragged_unpack: bit-exact copies of the valid part of every row (waveform: rows = 1, unit = hop;
    mel: rows = n_mels, unit = 1), zero tails on the way back, lengths clamped to [0, max_units]."""
    g = torch.Generator().manual_seed(rows * 1000 + unit)
    B = 7
    padded = torch.randn(B, rows, max_units * unit, generator=g).to(DEV)
    lens = torch.tensor([max_units, 0, 1, max_units // 2, max_units - 1, 3, max_units + 5])   # the last one is clamped
    eff = lens.clamp(0, max_units)
    sizes = eff * unit * rows
    offs = (torch.cumsum(sizes, 0) - sizes)
    if unit % 4 == 0:
        offs = offs                                   # (multiples of 4 elements: the 128-bit path)
    packed = torch.full((int(sizes.sum()) + 8,), 7.0, device=DEV)
    _ragged_copy(True, padded.reshape(B, -1) if rows == 1 else padded, packed, lens.to(DEV), offs.to(DEV), lens.tolist(), offs.tolist(),
                 rows, unit)
    ref = torch.cat([padded[b, :, : int(eff[b]) * unit].reshape(-1) for b in range(B)]).cpu()
    assert torch.equal(packed[: ref.numel()].cpu(), ref) and bool((packed[ref.numel():] == 7.0).all())
    back = torch.full_like(padded, 3.0)
    _ragged_copy(False, back.reshape(B, -1) if rows == 1 else back, packed, lens.to(DEV), offs.to(DEV), lens.tolist(), offs.tolist(),
                 rows, unit, zero_tail=True)
    for b in range(B):
        n = int(eff[b]) * unit
        assert torch.equal(back[b, :, :n], padded[b, :, :n]) and not back[b, :, n:].any()


@pytest.mark.parametrize("forced", [True, False])
def test_sharded_forward_single_gpu_matches_forward(forced):
    cfg = zo.ZeroVoxConfig.tiny()
    w = zo.make_weights(cfg, seed=1, dur_bias=float(np.log(4.0)))
    model = build_model(cfg, w, device=DEV)
    x = zo.make_inputs(cfg, 5, 11, 24, seed=7, ragged=True, dur_lo=0, dur_hi=5)
    with torch.no_grad():
        wav, mel, mel_len, logd = model(dict(x), force_duration=forced)
        # exact tails: the padded tensors of the plain call
        pw, pm, pl, pd = sharded_forward(model, dict(x), force_duration=forced, device=DEV, hop_length=cfg.hop_length,
                                         n_mels=cfg.n_mels, tails="padded")
        assert torch.equal(pw, wav) and torch.equal(pm, mel) and torch.equal(pl, mel_len) and torch.equal(pd, logd)
        # valid parts only, vocoded in 3 delivery groups, waveforms delivered to pinned host memory on the way
        host = torch.zeros(int(mel_len.sum()) * cfg.hop_length + 16).pin_memory()
        rb = sharded_forward(model, dict(x), force_duration=forced, device=DEV, hop_length=cfg.hop_length, n_mels=cfg.n_mels,
                             ragged=True, vocoder_groups=-3, host_out=host)
        torch.cuda.synchronize()
    o = 0
    for i, n in enumerate(mel_len.tolist()):
        assert torch.equal(rb.wav(i), wav[i, : n * cfg.hop_length]) and torch.equal(rb.mel(i), mel[i, :, :n])
        assert torch.equal(host[o: o + n * cfg.hop_length], wav[i, : n * cfg.hop_length].cpu())
        o += n * cfg.hop_length
    assert torch.equal(rb.log_duration(), logd) and rb.mel_len_host == mel_len.tolist()
    vw, vm, vl, vd = rb.padded()
    for i, n in enumerate(mel_len.tolist()):
        assert torch.equal(vw[i, : n * cfg.hop_length], wav[i, : n * cfg.hop_length]) and not vw[i, n * cfg.hop_length:].any()
        assert torch.equal(vm[i, :, :n], mel[i, :, :n]) and not vm[i, :, n:].any()


def test_shared_host_window_single_gpu():
    """SharedHostBatch / SharedHostBuffer with one rank: the window is page-locked (cudaHostRegister), the forward uploads
    from it and the waveforms land in it group by group."""
    from zerovox_b200.parallel import SharedHostBatch, SharedHostBuffer
    cfg = zo.ZeroVoxConfig.tiny()
    w = zo.make_weights(cfg, seed=1, dur_bias=float(np.log(4.0)))
    model = build_model(cfg, w, device=DEV)
    x = zo.make_inputs(cfg, 4, 9, 24, seed=3, ragged=True, dur_lo=1, dur_hi=5)
    with torch.no_grad():
        wav, mel, mel_len, logd = model(dict(x), force_duration=True)
        batch = SharedHostBatch(dict(x))
        assert batch.window.registered and batch.x["ref_mel"].is_pinned()
        win = SharedHostBuffer(4 * 4 * int(mel_len.max()) * cfg.hop_length)
        rb = sharded_forward(model, batch, force_duration=True, device=DEV, hop_length=cfg.hop_length, n_mels=cfg.n_mels,
                             ragged=True, vocoder_groups=2, host_out=win)
        torch.cuda.synchronize()
    for i, n in enumerate(mel_len.tolist()):
        assert torch.equal(rb.host_wav(i), wav[i, : n * cfg.hop_length].cpu())
        assert torch.equal(rb.wav(i), wav[i, : n * cfg.hop_length])          # one rank: also on the device
        assert torch.equal(rb.mel(i), mel[i, :, :n])
    win.close()
    batch.window.close()
