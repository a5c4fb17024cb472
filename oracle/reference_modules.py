"""The reference's OWN PyTorch modules as a CPU baseline / oracle anchor (test infrastructure, never shipped).

Two places hold the unmodified reference sources:
  * ``/root/reference`` — the read-only upstream tree, present in the build container only;
  * ``oracle/_ref/``    — a git-ignored copy of the six files of the path (``zerovox/tts/{model,fs2,hifigan,styletts,
                          ResNetSE34V2,symbols}.py``) written by ``oracle/build_ref.py`` next to a 10-line ``lightning``
                          stub; it travels to the GPU box with the snapshot, the sources never enter the history.
Only ``tests/``, ``__graft_entry__.smoke()``, ``oracle/make_goldens.py`` and ``bench.py``'s reference / ``cpu_baseline``
legs import this module.

``build_reference_model`` instantiates ``zerovox.tts.model.ZeroVox`` (model.py:158-254) with the kwargs mapping of
utils/train_tts.py:202-241 and loads reference-keyed weights; ``reference_forward`` is model.py:260-290 verbatim followed
by the intended HiFi-GAN tail ``wav = _meldec(mel.transpose(1, 2)).squeeze(1)`` (the reference's own eval tail,
model.py:298-304, is ParallelWaveGAN leftover code that raises with ``hifigan.Generator``).
"""
from __future__ import annotations

import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref")
REF_FILES = ("model.py", "fs2.py", "hifigan.py", "styletts.py", "ResNetSE34V2.py", "symbols.py")


def _install_lightning_stub():
    if "lightning" in sys.modules:
        return
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


def reference_root() -> str | None:
    """Directory to put on sys.path so that ``import zerovox.tts.model`` finds the unmodified reference."""
    env = os.environ.get("ZEROVOX_REFERENCE")
    for cand in (env, REF_COPY, "/root/reference"):
        if cand and os.path.exists(os.path.join(cand, "zerovox", "tts", "model.py")):
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


def import_reference():
    """Returns the reference's ``zerovox.tts`` modules (model, symbols, hifigan)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference modules not available: run `python oracle/build_ref.py` in the build container")
    _install_lightning_stub()
    if root not in sys.path:
        sys.path.insert(0, root)
    import zerovox.tts.model as model
    import zerovox.tts.symbols as symbols
    import zerovox.tts.hifigan as hifigan
    return model, symbols, hifigan


def build_reference_model(cfg, w: dict):
    """The reference's real ``ZeroVox`` + ``hifigan.Generator`` (weight norm removed) with ``w`` loaded, eval mode."""
    model, symbols, hifigan = import_reference()
    zv = model.ZeroVox(symbols=symbols.Symbols(cfg.phones, cfg.puncts), meldec_model=None,
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
    gen = hifigan.Generator(model.AttrDict(cfg.hifigan.as_json_dict())).eval()
    gen.remove_weight_norm()
    zv._meldec = gen
    missing, unexpected = zv.load_state_dict(w, strict=False)
    missing = [k for k in missing if "torchfb" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return zv.eval()


@torch.no_grad()
def reference_forward(zv, x, force_duration, style_embed=None):
    """model.py:260-290 verbatim, then the intended HiFi-GAN tail (module docstring).  ``style_embed`` replaces the
    speaker net's output (stage-wise parity tests feed the engine's own style vector).
    Returns (wav, mel [B, n_mels, L], mel_len, log_duration, pred, style)."""
    style = zv._spkemb(x["ref_mel"]) if style_embed is None else style_embed
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
