#!/usr/bin/env python
"""Pin the oracle against the reference's own modules and write tests/golden/*.npz.

Runs ONLY in the build container (needs /root/reference; never on the GPU box).
For each case it
  1. builds the reference's real nn.Modules (zerovox.tts.model.ZeroVox via a 10-line
     `lightning` stub, hifigan.Generator with remove_weight_norm()) and loads the seeded
     weights of oracle.zerovox_oracle.make_weights() into them,
  2. runs the reference path (ZeroVox sub-modules composed as model.py:260-290 + the intended
     HiFi-GAN tail; and ZeroVox.inference_ex itself for batch 1),
  3. asserts the functional restatement in oracle/zerovox_oracle.py agrees (fp32 round-off
     only; integer outputs exactly),
  4. stores inputs-by-seed + reference outputs as small fixtures.

Usage:  python oracle/make_goldens.py   (from the repo root)
"""
from __future__ import annotations

import dataclasses
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("ZEROVOX_REFERENCE", "/root/reference")

from oracle import zerovox_oracle as zo  # noqa: E402


def _install_lightning_stub():
    import torch.nn as nn

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    class LightningDataModule:
        pass

    m = types.ModuleType("lightning")
    m.LightningModule = LightningModule
    m.LightningDataModule = LightningDataModule
    sys.modules["lightning"] = m


def build_reference_model(cfg: zo.ZeroVoxConfig, w: dict):
    _install_lightning_stub()
    sys.path.insert(0, REF)
    from zerovox.tts.model import ZeroVox, AttrDict
    from zerovox.tts.symbols import Symbols
    from zerovox.tts.hifigan import Generator

    zv = ZeroVox(symbols=Symbols(cfg.phones, cfg.puncts), meldec_model=None,
                 sampling_rate=cfg.sampling_rate, hop_length=cfg.hop_length, n_mels=cfg.n_mels,
                 lr=1e-4, weight_decay=0.0, max_epochs=1, warmup_epochs=1, betas=(0.0, 0.99), eps=1e-9,
                 embed_dim=cfg.emb_dim, punct_embed_dim=cfg.punct_emb_dim, dpe_embed_dim=32, emb_reduction=1,
                 max_mel_len=cfg.max_mel_len, max_txt_len=cfg.max_txt_len,
                 fs2enc_layer=cfg.enc_layers, fs2enc_head=cfg.enc_heads, fs2enc_dropout=0.2,
                 vp_filter_size=cfg.vp_filter_size, vp_kernel_size=cfg.vp_kernel_size, vp_dropout=0.5,
                 ve_n_bins=cfg.ve_n_bins,
                 resnet_layers=list(cfg.resnet_layers), resnet_num_filters=list(cfg.resnet_num_filters),
                 resnet_encoder_type=cfg.resnet_encoder_type,
                 decoder_kind=cfg.decoder_kind, decoder_n_layers=cfg.dec_layers, decoder_n_head=cfg.dec_heads,
                 decoder_conv_filter_size=cfg.conv_filter_size,
                 decoder_conv_kernel_size=list(cfg.conv_kernel_size),
                 decoder_dropout=0.2, decoder_scln=cfg.dec_scln)
    gen = Generator(AttrDict(cfg.hifigan.as_json_dict())).eval()
    gen.remove_weight_norm()
    zv._meldec = gen
    missing, unexpected = zv.load_state_dict(w, strict=False)
    missing = [k for k in missing if "torchfb" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return zv.eval()


@torch.no_grad()
def reference_forward(zv, x, force_duration):
    """model.py:260-290 verbatim, then the intended HiFi-GAN tail (see oracle docstring)."""
    style = zv._spkemb(x["ref_mel"])
    pred = zv._phoneme_encoder(x, style_embed=style, train=False, force_duration=force_duration)
    mask = pred["masks"]
    if mask is None:
        L = pred["features"].shape[1]
        dec_mask = ~(torch.arange(L).expand(len(pred["mel_len"]), L) < pred["mel_len"].unsqueeze(1))
    else:
        dec_mask = mask[:, :, 0]
    mel, _ = zv._mel_decoder(pred["features"], dec_mask, spk_emb=style)
    if mask is not None and mel.size(0) > 1:
        mel = mel.masked_fill(mask[:, :, : mel.shape[-1]], 0)
    wav = zv._meldec(mel.transpose(1, 2)).squeeze(1)
    return wav, mel.transpose(1, 2), pred["mel_len"], pred["log_duration"], pred, style


def check(name, a, b, atol, rtol=0.0):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    err = (a - b).abs().max().item() if a.numel() else 0.0
    ref = b.abs().max().item() if b.numel() else 0.0
    ok = err <= atol + rtol * ref
    print(f"    {name:22s} max|diff|={err:.3e}  max|ref|={ref:.3e}  {'ok' if ok else 'FAIL'}")
    assert ok, name


CASES = [
    # name, config, seed_w, B, T, T_ref, ragged, force_duration, dur range
    ("tiny_forced", zo.ZeroVoxConfig.tiny(), 1, 3, 11, 24, True, True, (0, 5)),
    ("tiny_predicted", zo.ZeroVoxConfig.tiny(), 2, 2, 9, 32, True, False, (1, 4)),
    ("tiny_longform", zo.ZeroVoxConfig.tiny(), 3, 1, 30, 16, False, True, (2, 4)),  # T>max_txt_len, L>max_mel_len
    ("medium_forced", zo.ZeroVoxConfig(), 0, 2, 12, 48, True, True, (2, 7)),
    ("medium_predicted", zo.ZeroVoxConfig(), 0, 2, 10, 40, False, False, (2, 7)),
    # decoder_kind="styletts" (configs/tts_medium_styledec.yaml: the shipped default models, BASELINE config 1)
    ("tiny_styledec", dataclasses.replace(zo.ZeroVoxConfig.tiny(), decoder_kind="styletts"), 4, 3, 10, 24, True, True, (1, 5)),
    ("medium_styledec", dataclasses.replace(zo.ZeroVoxConfig(), decoder_kind="styletts"), 0, 2, 11, 40, True, False, (2, 7)),
]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, cfg, seed_w, B, T, T_ref, ragged, force, (dlo, dhi) in CASES:
        print(f"[{name}] B={B} T={T} T_ref={T_ref} ragged={ragged} force_duration={force}")
        w = zo.make_weights(cfg, seed=seed_w, dur_bias=np.log(4.0))
        x = zo.make_inputs(cfg, B, T, T_ref, seed=7, ragged=ragged, dur_lo=dlo, dur_hi=dhi)
        zv = build_reference_model(cfg, w)
        wav, mel, mel_len, logd, pred, style = reference_forward(zv, dict(x), force)
        with torch.no_grad():
            owav, omel, omel_len, ologd, st = zo.zerovox_forward(cfg, w, dict(x), force_duration=force)
        # restatement == reference modules
        check("style_embed", st["style_embed"], style, 2e-6)
        check("pitch", st["pitch"], pred["pitch"], 2e-5)
        check("energy", st["energy"], pred["energy"], 2e-5)
        check("log_duration", ologd, logd, 2e-5)
        assert torch.equal(omel_len, mel_len), (omel_len, mel_len)
        check("features", st["features"], pred["features"], 2e-5)
        check("mel", omel, mel, 1e-4, 1e-5)
        check("wav", owav, wav, 1e-4)
        if force:
            assert torch.equal(mel_len, x["duration"].clamp(min=0).sum(1).long())  # export_hifigan.py:125-128
        assert wav.shape[1] == mel.shape[2] * cfg.hop_length
        gold = {
            "seed_w": seed_w, "seed_x": 7, "B": B, "T": T, "T_ref": T_ref, "ragged": ragged, "force": force,
            "dur_lo": dlo, "dur_hi": dhi, "dur_bias": np.log(4.0),
            "style_embed": style.numpy(), "pitch": pred["pitch"].numpy(), "energy": pred["energy"].numpy(),
            "log_duration": logd.numpy(), "mel_len": mel_len.numpy(),
            "src_index": st["_src_index"], "pitch_bucket": st["_pitch_bucket"].numpy().astype(np.int16),
            "energy_bucket": st["_energy_bucket"].numpy().astype(np.int16),
            "duration_rounded": st["_duration_rounded"].numpy().astype(np.int32),
            "mel": mel.numpy(), "wav": wav.numpy(),
        }
        # batch-1 inference_ex through the reference's own ZeroVox.inference_ex (model.py:308-347)
        x1 = {k: v[:1] for k, v in x.items()}
        if "phoneme_mask" in x1:
            x1.pop("phoneme_mask")  # tts_ex passes no mask (synthesize.py:228-231)
        zv._min_mel_len = 40 if cfg.max_mel_len < 100 else 100
        m0 = zv._min_mel_len
        with torch.no_grad():
            rwav, rlen, rlogd, rmel = zv.inference_ex(dict(x1), style_embed=style[:1], force_duration=force)
            iwav, ilen, ilogd, imel, mml = zo.zerovox_inference_ex(cfg, w, dict(x1), style[:1], force_duration=force,
                                                                   min_mel_len=m0)
        assert ilen == rlen and mml == zv._min_mel_len
        check("inference_ex.wav", iwav, rwav, 1e-4)
        check("inference_ex.mel", imel, rmel, 1e-4, 1e-5)
        check("inference_ex.logd", ilogd, rlogd, 2e-5)
        gold.update({"ix_min_mel_len": m0, "ix_wav": rwav.numpy(), "ix_mel": rmel.numpy(), "ix_mel_len": rlen,
                     "ix_log_duration": rlogd.numpy()})
        path = os.path.join(out_dir, f"{name}.npz")
        np.savez_compressed(path, **gold)
        print(f"    -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)  mel_len={mel_len.tolist()}")

    # vocoder-only goldens for V1 / V3 topologies (config 3) against hifigan.Generator
    _install_lightning_stub()
    sys.path.insert(0, REF)
    from zerovox.tts.hifigan import Generator
    from zerovox.tts.model import AttrDict
    for vname, h in (("v1", zo.HifiGanConfig.v1()), ("v2", zo.HifiGanConfig.v2()), ("v3", zo.HifiGanConfig.v3())):
        g = torch.Generator().manual_seed(11)
        hw = zo.make_hifigan_weights(h, g)
        mel = torch.randn((2, 80, 9), generator=g)
        gen = Generator(AttrDict(h.as_json_dict())).eval()
        gen.remove_weight_norm()
        gen.load_state_dict(hw)
        with torch.no_grad():
            ref = gen(mel)
            ours = zo.hifigan_generator(h, hw, mel, prefix="")
        print(f"[hifigan_{vname}]")
        check("wav", ours, ref, 2e-5)
        np.savez_compressed(os.path.join(out_dir, f"hifigan_{vname}.npz"), seed=11, mel=mel.numpy(), wav=ref.numpy())


if __name__ == "__main__":
    main()
