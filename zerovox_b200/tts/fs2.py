"""FS2Encoder / FS2Decoder with the reference's constructor arguments, forward() signatures and state_dict keys
(zerovox/tts/fs2.py:232-315, 697-775).  The classes below only *hold parameters*; the eval-mode arithmetic runs in
the CUDA engine (zvx_encode / zvx_length_regulate / zvx_decode)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ._context import EngineModuleMixin


def _sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """Value-identical to fs2.py:17-37 (float64 angles -> sin/cos -> fp32)."""
    ang = np.arange(n_position, dtype=np.float64)[:, None] / np.power(10000, 2 * (np.arange(d_hid) // 2) / d_hid)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(ang.astype(np.float32))


class _Holder(nn.Module):
    """A parameter container; forward lives in the engine."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: call the owning FS2Encoder / FS2Decoder instead")


class _SCLNParams(_Holder):          # keys: affine_layer.linear.weight   (fs2.py:63-106)
    def __init__(self, s_size, hidden):
        super().__init__()
        self.affine_layer = _Holder()
        self.affine_layer.linear = nn.Linear(s_size, 2 * hidden, bias=False)
        nn.init.xavier_uniform_(self.affine_layer.linear.weight)


class _AttnParams(_Holder):          # keys: w_qs, w_ks, w_vs, fc, layer_norm   (fs2.py:108-131)
    def __init__(self, d_model, spk, scln):
        super().__init__()
        self.w_qs, self.w_ks, self.w_vs = (nn.Linear(d_model, d_model) for _ in range(3))
        self.layer_norm = _SCLNParams(spk, d_model) if scln else nn.LayerNorm(d_model)
        self.fc = nn.Linear(d_model, d_model)


class _FFNParams(_Holder):           # keys: w_1, w_2, layer_norm   (fs2.py:166-194)
    def __init__(self, d_in, d_hid, kernel_size, spk, scln):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, kernel_size[0], padding=(kernel_size[0] - 1) // 2)
        self.w_2 = nn.Conv1d(d_hid, d_in, kernel_size[1], padding=(kernel_size[1] - 1) // 2)
        self.layer_norm = _SCLNParams(spk, d_in) if scln else nn.LayerNorm(d_in)


class _FFTBlockParams(_Holder):      # keys: slf_attn.*, pos_ffn.*   (fs2.py:211-219)
    def __init__(self, d_model, d_inner, kernel_size, spk, scln):
        super().__init__()
        self.slf_attn = _AttnParams(d_model, spk, scln)
        self.pos_ffn = _FFNParams(d_model, d_inner, kernel_size, spk, scln)


class _ConvParams(_Holder):          # key: conv   (fs2.py:461-497)
    def __init__(self, cin, cout, k, padding):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, k, padding=padding)


class _VarPredictorParams(_Holder):  # keys: conv_layer.{conv1d_1.conv,layer_norm_1,conv1d_2.conv,layer_norm_2}, linear_layer
    def __init__(self, emb, filt, k):
        super().__init__()
        cl = _Holder()
        cl.conv1d_1 = _ConvParams(emb, filt, k, (k - 1) // 2)
        cl.layer_norm_1 = nn.LayerNorm(filt)
        cl.conv1d_2 = _ConvParams(filt, filt, k, 1)
        cl.layer_norm_2 = nn.LayerNorm(filt)
        self.conv_layer = cl
        self.linear_layer = nn.Linear(filt, 1)


class FS2Encoder(EngineModuleMixin, nn.Module):
    """zerovox/tts/fs2.py:697-775."""
    _role = "encoder"

    def __init__(self, symbols, max_txt_len, embed_dim, encoder_layer, encoder_head, conv_filter_size,
                 conv_kernel_size, encoder_dropout, punct_embed_dim, vp_filter_size, vp_kernel_size, vp_dropout,
                 ve_n_bins):
        super().__init__()
        hidden = embed_dim + punct_embed_dim
        self._hp = dict(num_phones=symbols.num_phones, num_puncts=symbols.num_puncts, emb_dim=embed_dim,
                        punct_emb_dim=punct_embed_dim, max_txt_len=max_txt_len, enc_layers=encoder_layer,
                        enc_heads=encoder_head, conv_filter_size=conv_filter_size,
                        conv_kernel_size=tuple(conv_kernel_size), vp_filter_size=vp_filter_size,
                        vp_kernel_size=vp_kernel_size, ve_n_bins=ve_n_bins)
        enc = _Holder()
        enc.src_word_emb = nn.Embedding(symbols.num_phones + 1, embed_dim, padding_idx=0)
        enc.punct_embed = nn.Embedding(symbols.num_puncts + 1, punct_embed_dim, padding_idx=0)
        enc.position_enc = nn.Parameter(_sinusoid_table(max_txt_len + 1, hidden).unsqueeze(0), requires_grad=False)
        enc.layer_stack = nn.ModuleList(
            _FFTBlockParams(hidden, conv_filter_size, conv_kernel_size, 0, False) for _ in range(encoder_layer))
        self._encoder = enc
        va = _Holder()
        va.duration_predictor = _VarPredictorParams(hidden, vp_filter_size, vp_kernel_size)
        va.pitch_predictor = _VarPredictorParams(hidden, vp_filter_size, vp_kernel_size)
        va.energy_predictor = _VarPredictorParams(hidden, vp_filter_size, vp_kernel_size)
        va.pitch_embedding = nn.Embedding(ve_n_bins, hidden)
        va.energy_embedding = nn.Embedding(ve_n_bins, hidden)
        self._variance_adaptor = va
        self._init_engine_binding()

    def _fill_config(self, cfg):
        for k, v in self._hp.items():
            setattr(cfg, k, v)

    def forward(self, x, style_embed, train=False, force_duration=False):
        """Same contract as fs2.py:732-775: returns the dict pitch/energy/log_duration/mel_len/features/masks."""
        if train:
            raise NotImplementedError("FS2Encoder(train=True): training is outside the zerovox_b200 hot path")
        eng = self._engine()
        dev = eng.device
        phoneme = x["phoneme"].to(dev)
        puncts = x["puncts"].to(dev)
        mask = x["phoneme_mask"].to(dev) if "phoneme_mask" in x else None
        forced = x["duration"].to(dev) if force_duration else None
        r = eng.encode(phoneme, puncts, style_embed.to(dev), mask, forced, need_lengths=True)
        feats = eng.length_regulate(r["xprime"], r["duration_rounded"], r["L_max"])
        masks = None
        if not force_duration:  # mel mask only exists when durations were predicted (fs2.py:683, 772)
            mel_mask = torch.arange(r["L_max"], device=dev)[None, :] >= r["mel_len"][:, None]
            masks = mel_mask.unsqueeze(2).expand(-1, -1, feats.shape[2])
        return {"pitch": r["pitch"], "energy": r["energy"], "log_duration": r["log_duration"],
                "mel_len": r["mel_len"], "features": feats, "masks": masks,
                "_duration_rounded": r["duration_rounded"], "_mel_len_host": r["mel_len_host"]}


class FS2Decoder(EngineModuleMixin, nn.Module):
    """zerovox/tts/fs2.py:232-315."""
    _role = "decoder"

    def __init__(self, dec_max_seq_len, dec_hidden, dec_n_layers, dec_n_head, dec_conv_filter_size,
                 dec_conv_kernel_size, dec_dropout, dec_scln, n_mel_channels, spk_emb_size):
        super().__init__()
        if spk_emb_size != dec_hidden and dec_scln:
            raise ValueError("zerovox_b200: SCLN expects spk_emb_size == dec_hidden (as in model.py:216-236)")
        self._hp = dict(max_mel_len=dec_max_seq_len, dec_layers=dec_n_layers, dec_heads=dec_n_head,
                        conv_filter_size=dec_conv_filter_size, conv_kernel_size=tuple(dec_conv_kernel_size),
                        dec_scln=bool(dec_scln), n_mels=n_mel_channels, decoder_kind="fastspeech2")
        self._hidden = dec_hidden
        self.max_seq_len = dec_max_seq_len
        self.d_model = dec_hidden
        self.position_enc = nn.Parameter(_sinusoid_table(dec_max_seq_len + 1, dec_hidden).unsqueeze(0),
                                         requires_grad=False)
        self.layer_stack = nn.ModuleList(
            _FFTBlockParams(dec_hidden, dec_conv_filter_size, dec_conv_kernel_size, spk_emb_size, dec_scln)
            for _ in range(dec_n_layers))
        self.mel_linear = nn.Linear(dec_hidden, n_mel_channels)
        self._init_engine_binding()

    def _fill_config(self, cfg):
        for k, v in self._hp.items():
            setattr(cfg, k, v)
        if cfg.hidden != self._hidden:  # stand-alone decoder: make hidden consistent
            cfg.emb_dim, cfg.punct_emb_dim = self._hidden - 16, 16

    def forward(self, enc_seq, mask, spk_emb, return_attns=False):
        """fs2.py:281-315: (enc_seq [B,L,H], mask bool [B,L], spk_emb [B,1,H]) -> (mel [B,L,n_mels], mask)."""
        eng = self._engine()
        mel, _ = eng.decode(enc_seq.to(eng.device), spk_emb.to(eng.device), mask=mask.to(eng.device),
                            want_blc=True, want_bcl=False)
        return mel, mask
