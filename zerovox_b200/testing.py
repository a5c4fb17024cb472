"""Helpers shared by tests / smoke / bench: build the mirror model from an oracle-style config object and a
reference-keyed weight dict.  (No oracle import here: callers pass plain objects.)"""
from __future__ import annotations

import torch

from .tts.model import ZeroVox, AttrDict
from .tts.hifigan import Generator
from .tts.symbols import Symbols


def zerovox_kwargs(cfg) -> dict:
    """ZeroVox.__init__ kwargs from a config object with the oracle's field names (kwargs mapping of
    utils/train_tts.py:202-241)."""
    return dict(symbols=Symbols(cfg.phones, cfg.puncts), meldec_model=None, sampling_rate=cfg.sampling_rate,
                hop_length=cfg.hop_length, n_mels=cfg.n_mels, lr=1e-4, weight_decay=0.0, max_epochs=1,
                warmup_epochs=1, betas=(0.0, 0.99), eps=1e-9, embed_dim=cfg.emb_dim,
                punct_embed_dim=cfg.punct_emb_dim, dpe_embed_dim=32, emb_reduction=1, max_mel_len=cfg.max_mel_len,
                max_txt_len=cfg.max_txt_len, fs2enc_layer=cfg.enc_layers, fs2enc_head=cfg.enc_heads,
                fs2enc_dropout=0.2, vp_filter_size=cfg.vp_filter_size, vp_kernel_size=cfg.vp_kernel_size,
                vp_dropout=0.5, ve_n_bins=cfg.ve_n_bins, resnet_layers=list(cfg.resnet_layers),
                resnet_num_filters=list(cfg.resnet_num_filters), resnet_encoder_type=cfg.resnet_encoder_type,
                decoder_kind=cfg.decoder_kind, decoder_n_layers=cfg.dec_layers, decoder_n_head=cfg.dec_heads,
                decoder_conv_filter_size=cfg.conv_filter_size, decoder_conv_kernel_size=list(cfg.conv_kernel_size),
                decoder_dropout=0.2, decoder_scln=cfg.dec_scln)


def build_generator(hcfg, weights: dict | None = None, prefix: str = "_meldec.") -> Generator:
    gen = Generator(AttrDict(hcfg.as_json_dict())).eval()
    gen.remove_weight_norm()
    if weights is not None:
        gen.load_state_dict({k[len(prefix):]: v for k, v in weights.items() if k.startswith(prefix)})
    return gen


def build_model(cfg, weights: dict, device=None, tensor_core_policy: int = 1) -> ZeroVox:
    """Mirror ZeroVox + Generator with ``weights`` (reference state_dict keys) loaded, in eval mode."""
    zv = ZeroVox(**zerovox_kwargs(cfg))
    zv._meldec = build_generator(cfg.hifigan)
    missing, unexpected = zv.load_state_dict(weights, strict=False)
    missing = [k for k in missing if "torchfb" not in k]
    if missing or unexpected:
        raise RuntimeError(f"state_dict mismatch: missing={missing[:5]} unexpected={list(unexpected)[:5]}")
    zv._shared_ctx.tensor_core_policy = tensor_core_policy
    zv.eval()
    if device is not None:
        zv.to(device)
    return zv
