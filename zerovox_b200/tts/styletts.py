"""StyleTTS mel decoder with the reference's interface (zerovox/tts/styletts.py:142-205): same constructor, same
``state_dict`` keys (convolutions stay in weight-norm form — ``weight_g`` / ``weight_v`` — because the reference never
removes it there, styletts.py:28-34, 113-118), ``forward(enc_seq, mask, spk_emb) -> (mel, None)``.  The arithmetic
runs in the CUDA engine (zvx_decode with decoder_kind = styletts)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ._context import EngineModuleMixin


class _WNConv(nn.Module):
    """Parameter holder of weight_norm(nn.Conv1d(cin, cout, k)) — keys weight_g [cout,1,1], weight_v [cout,cin,k], bias."""

    def __init__(self, cin, cout, k, bias=True):
        super().__init__()
        v = torch.empty(cout, cin, k)
        nn.init.kaiming_uniform_(v, a=5 ** 0.5)
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).view(-1, 1, 1).clone())
        self.weight_v = nn.Parameter(v)
        if bias:
            self.bias = nn.Parameter(torch.zeros(cout))


class _Affine(nn.Module):
    """InstanceNorm1d(affine=True) parameters: weight, bias (no running stats, as in the reference)."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


class _AdaIN(nn.Module):
    def __init__(self, style_dim, c):
        super().__init__()
        self.fc = nn.Linear(style_dim, 2 * c)


class _ResBlk(nn.Module):        # styletts.py:11-69 (normalize=True, downsample='none')
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = _WNConv(cin, cin, 3)
        self.conv2 = _WNConv(cin, cout, 3)
        self.norm1 = _Affine(cin)
        self.norm2 = _Affine(cin)
        if cin != cout:
            self.conv1x1 = _WNConv(cin, cout, 1, bias=False)


class _AdainResBlk(nn.Module):   # styletts.py:95-139
    def __init__(self, cin, cout, style_dim):
        super().__init__()
        self.conv1 = _WNConv(cin, cout, 3)
        self.conv2 = _WNConv(cout, cout, 3)
        self.norm1 = _AdaIN(style_dim, cin)
        self.norm2 = _AdaIN(style_dim, cout)
        if cin != cout:
            self.conv1x1 = _WNConv(cin, cout, 1, bias=False)


class StyleTTSDecoder(EngineModuleMixin, nn.Module):
    _role = "decoder"

    def __init__(self, dim_in, style_dim, residual_dim, dim_out):
        super().__init__()
        if style_dim != dim_in or residual_dim != 64:
            raise ValueError("zerovox_b200: StyleTTSDecoder is built as in model.py:238-242 (style_dim == dim_in, "
                             "residual_dim == 64)")
        self._hp = dict(decoder_kind="styletts", n_mels=dim_out)
        self._hidden = dim_in
        bn = dim_in * 2
        self.bottleneck_dim = bn
        self.encode = nn.Sequential(_ResBlk(dim_in, bn), _ResBlk(bn, bn))
        self.decode = nn.ModuleList([_AdainResBlk(bn + residual_dim, bn, style_dim),
                                     _AdainResBlk(bn + residual_dim, bn, style_dim),
                                     _AdainResBlk(bn + residual_dim, dim_in, style_dim),
                                     _AdainResBlk(dim_in, dim_in, style_dim),
                                     _AdainResBlk(dim_in, dim_in, style_dim)])
        self.asr_res = nn.Sequential(_WNConv(dim_in, residual_dim, 1), _Affine(residual_dim))
        self.to_out = nn.Sequential(_WNConv(dim_in, dim_out, 1))
        self._init_engine_binding()

    def _fill_config(self, cfg):
        for k, v in self._hp.items():
            setattr(cfg, k, v)
        if cfg.hidden != self._hidden:  # stand-alone decoder: make hidden consistent
            cfg.emb_dim, cfg.punct_emb_dim = self._hidden - 16, 16

    def forward(self, enc_seq, mask, spk_emb):
        """styletts.py:181-205: (enc_seq [B,L,H], mask (ignored, as in the reference), spk_emb [B,1,H]) -> (mel, None)."""
        eng = self._engine()
        B, L, _ = enc_seq.shape
        lens = torch.full((B,), L, dtype=torch.int64, device=eng.device)
        mel, _ = eng.decode(enc_seq.to(eng.device), spk_emb.to(eng.device), mel_len=lens, want_blc=True, want_bcl=False)
        return mel, None
