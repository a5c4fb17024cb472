"""Synthetic workloads: model configurations (configs/tts_medium.yaml + upstream HiFi-GAN config_v{1,2,3}.json),
seeded random weights keyed like the reference state_dict, and seeded input batches.

Pure Python / torch-CPU data generation shared by tests, smoke(), bench.py and the oracle; nothing here computes
the forward path.  (There is no network for checkpoints or datasets, so every parity and bench input is synthetic.)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from .tts.fs2 import _sinusoid_table


# ----------------------------------------------------------------------------
# configuration (configs/tts_medium.yaml:3-51 + upstream HiFi-GAN config_v{1,2,3}.json)
# ----------------------------------------------------------------------------
@dataclass
class HifiGanConfig:
    resblock: str = "1"
    upsample_rates: tuple = (8, 8, 2, 2)
    upsample_kernel_sizes: tuple = (16, 16, 4, 4)
    upsample_initial_channel: int = 128
    resblock_kernel_sizes: tuple = (3, 7, 11)
    resblock_dilation_sizes: tuple = ((1, 3, 5), (1, 3, 5), (1, 3, 5))

    @staticmethod
    def v1():
        return HifiGanConfig(upsample_initial_channel=512)

    @staticmethod
    def v2():
        return HifiGanConfig(upsample_initial_channel=128)

    @staticmethod
    def v3():
        return HifiGanConfig(resblock="2", upsample_rates=(8, 8, 4), upsample_kernel_sizes=(16, 16, 8),
                             upsample_initial_channel=256, resblock_kernel_sizes=(3, 5, 7),
                             resblock_dilation_sizes=((1, 2), (2, 6), (3, 12)))

    def as_json_dict(self):
        return {"resblock": self.resblock, "upsample_rates": list(self.upsample_rates),
                "upsample_kernel_sizes": list(self.upsample_kernel_sizes),
                "upsample_initial_channel": self.upsample_initial_channel,
                "resblock_kernel_sizes": list(self.resblock_kernel_sizes),
                "resblock_dilation_sizes": [list(d) for d in self.resblock_dilation_sizes]}


@dataclass
class ZeroVoxConfig:
    """tts_medium.yaml; kwargs mapping of utils/train_tts.py:202-241."""
    phones: str = "'-abcdefghijklmnopqrstuvwxyz"
    puncts: str = " ,.;:-!?\""
    emb_dim: int = 512
    punct_emb_dim: int = 16
    max_txt_len: int = 512
    max_mel_len: int = 1750
    enc_layers: int = 4
    enc_heads: int = 2
    vp_filter_size: int = 256
    vp_kernel_size: int = 3
    ve_n_bins: int = 256
    decoder_kind: str = "fastspeech2"
    dec_layers: int = 6
    dec_heads: int = 2
    conv_filter_size: int = 1024
    conv_kernel_size: tuple = (9, 1)
    dec_scln: bool = True
    resnet_layers: tuple = (3, 4, 6, 3)
    resnet_num_filters: tuple = (32, 64, 128, 256)
    resnet_encoder_type: str = "ASP"
    n_mels: int = 80
    sampling_rate: int = 22050
    hop_length: int = 256
    hifigan: HifiGanConfig = field(default_factory=HifiGanConfig.v2)

    @property
    def hidden(self):
        return self.emb_dim + self.punct_emb_dim

    @property
    def num_phones(self):  # symbols.py:35-37
        return len(self.phones)

    @property
    def num_puncts(self):  # symbols.py:47-49 (includes _NP_)
        return len(self.puncts) + 1

    @staticmethod
    def tiny():
        """Small dims for fast CPU tests (same code path, every feature on)."""
        return ZeroVoxConfig(emb_dim=80, punct_emb_dim=16, max_txt_len=24, max_mel_len=60,
                             enc_layers=2, dec_layers=2, vp_filter_size=32, conv_filter_size=96,
                             resnet_layers=(1, 1, 1, 1), resnet_num_filters=(8, 8, 16, 16),
                             hifigan=HifiGanConfig(upsample_initial_channel=32))


# ----------------------------------------------------------------------------
# seeded weights, keyed like the reference state_dict
# ----------------------------------------------------------------------------
def _fan_in_normal(g, shape, fan_in, gain=1.0):
    return torch.randn(shape, generator=g, dtype=torch.float32) * (gain / math.sqrt(fan_in))


def make_weights(cfg: ZeroVoxConfig, seed: int = 0, dur_bias: float | None = None) -> dict:
    """Deterministic random weights with O(1) activations everywhere.

    Not the reference's init (that needs the reference importable, and its
    HiFi-GAN init N(0, 0.01) gives ~0 output, hifigan.py:17-20); any fp32 values
    are valid parity inputs.  ``dur_bias`` sets duration_predictor.linear bias
    (log(7) gives ~6 frames/phoneme with predicted durations).
    """
    g = torch.Generator().manual_seed(seed)
    w = {}
    H, DI = cfg.hidden, cfg.conv_filter_size
    k1, k2 = cfg.conv_kernel_size

    def fft_stack(prefix, n_layers, scln):
        for i in range(n_layers):
            p = f"{prefix}.layer_stack.{i}"
            for nm in ("w_qs", "w_ks", "w_vs", "fc"):
                w[f"{p}.slf_attn.{nm}.weight"] = _fan_in_normal(g, (H, H), H)
                w[f"{p}.slf_attn.{nm}.bias"] = _fan_in_normal(g, (H,), 16.0)
            w[f"{p}.pos_ffn.w_1.weight"] = _fan_in_normal(g, (DI, H, k1), H * k1, 1.4)
            w[f"{p}.pos_ffn.w_1.bias"] = _fan_in_normal(g, (DI,), 16.0)
            w[f"{p}.pos_ffn.w_2.weight"] = _fan_in_normal(g, (H, DI, k2), DI * k2, 1.4)
            w[f"{p}.pos_ffn.w_2.bias"] = _fan_in_normal(g, (H,), 16.0)
            for ln in ("slf_attn", "pos_ffn"):
                if scln:
                    # bias rows first, gain rows second (fs2.py:85); |style| = 1, so unit-variance
                    # rows give O(1) random-sign gains and biases (a non-degenerate SCLN)
                    aff = torch.randn((2 * H, H), generator=g)
                    w[f"{p}.{ln}.layer_norm.affine_layer.linear.weight"] = aff
                else:
                    w[f"{p}.{ln}.layer_norm.weight"] = 1.0 + 0.1 * torch.randn((H,), generator=g)
                    w[f"{p}.{ln}.layer_norm.bias"] = 0.1 * torch.randn((H,), generator=g)

    # encoder (fs2.py:350-368)
    e = "_phoneme_encoder._encoder"
    w[f"{e}.position_enc"] = _sinusoid_table(cfg.max_txt_len + 1, H).unsqueeze(0)
    emb = torch.randn((cfg.num_phones + 1, cfg.emb_dim), generator=g)
    emb[0] = 0.0  # padding_idx=0 (fs2.py:350)
    w[f"{e}.src_word_emb.weight"] = emb
    pemb = torch.randn((cfg.num_puncts + 1, cfg.punct_emb_dim), generator=g)
    pemb[0] = 0.0
    w[f"{e}.punct_embed.weight"] = pemb
    fft_stack(e, cfg.enc_layers, scln=False)

    # variance adaptor (fs2.py:586-624)
    va = "_phoneme_encoder._variance_adaptor"
    F_, K = cfg.vp_filter_size, cfg.vp_kernel_size
    for nm in ("duration", "pitch", "energy"):
        p = f"{va}.{nm}_predictor"
        w[f"{p}.conv_layer.conv1d_1.conv.weight"] = _fan_in_normal(g, (F_, H, K), H * K, 1.4)
        w[f"{p}.conv_layer.conv1d_1.conv.bias"] = _fan_in_normal(g, (F_,), 16.0)
        w[f"{p}.conv_layer.layer_norm_1.weight"] = 1.0 + 0.1 * torch.randn((F_,), generator=g)
        w[f"{p}.conv_layer.layer_norm_1.bias"] = 0.1 * torch.randn((F_,), generator=g)
        w[f"{p}.conv_layer.conv1d_2.conv.weight"] = _fan_in_normal(g, (F_, F_, K), F_ * K, 1.4)
        w[f"{p}.conv_layer.conv1d_2.conv.bias"] = _fan_in_normal(g, (F_,), 16.0)
        w[f"{p}.conv_layer.layer_norm_2.weight"] = 1.0 + 0.1 * torch.randn((F_,), generator=g)
        w[f"{p}.conv_layer.layer_norm_2.bias"] = 0.1 * torch.randn((F_,), generator=g)
        # pitch/energy predictions should span [0,1] so that many buckets are hit
        gain = 0.35 if nm != "duration" else 0.5
        w[f"{p}.linear_layer.weight"] = _fan_in_normal(g, (1, F_), F_, gain)
        b = 0.5 if nm != "duration" else (dur_bias if dur_bias is not None else math.log(7.0))
        w[f"{p}.linear_layer.bias"] = torch.full((1,), float(b))
    w[f"{va}.pitch_embedding.weight"] = 0.5 * torch.randn((cfg.ve_n_bins, H), generator=g)
    w[f"{va}.energy_embedding.weight"] = 0.5 * torch.randn((cfg.ve_n_bins, H), generator=g)

    # speaker net (ResNetSE34V2.py:101-155)
    s = "_spkemb"
    nf = cfg.resnet_num_filters
    w[f"{s}.conv1.weight"] = _fan_in_normal(g, (nf[0], 1, 3, 3), 9, 1.4)
    w[f"{s}.conv1.bias"] = 0.1 * torch.randn((nf[0],), generator=g)

    def bn(prefix, c):
        w[f"{prefix}.weight"] = 1.0 + 0.1 * torch.randn((c,), generator=g)
        w[f"{prefix}.bias"] = 0.1 * torch.randn((c,), generator=g)
        w[f"{prefix}.running_mean"] = 0.1 * torch.randn((c,), generator=g)
        w[f"{prefix}.running_var"] = 0.5 + torch.rand((c,), generator=g)
        w[f"{prefix}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    bn(f"{s}.bn1", nf[0])
    inpl = nf[0]
    for li, (planes, nblocks) in enumerate(zip(nf, cfg.resnet_layers), start=1):
        for bi in range(nblocks):
            p = f"{s}.layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            w[f"{p}.conv1.weight"] = _fan_in_normal(g, (planes, inpl, 3, 3), inpl * 9, 1.4)
            bn(f"{p}.bn1", planes)
            w[f"{p}.conv2.weight"] = _fan_in_normal(g, (planes, planes, 3, 3), planes * 9, 1.0)
            bn(f"{p}.bn2", planes)
            r = planes // 8
            w[f"{p}.se.fc.0.weight"] = _fan_in_normal(g, (r, planes), planes)
            w[f"{p}.se.fc.0.bias"] = 0.1 * torch.randn((r,), generator=g)
            w[f"{p}.se.fc.2.weight"] = _fan_in_normal(g, (planes, r), r)
            w[f"{p}.se.fc.2.bias"] = 0.1 * torch.randn((planes,), generator=g)
            if stride != 1 or inpl != planes:
                w[f"{p}.downsample.0.weight"] = _fan_in_normal(g, (planes, inpl, 1, 1), inpl)
                bn(f"{p}.downsample.1", planes)
            inpl = planes
    D = nf[3] * (cfg.n_mels // 8)
    w[f"{s}.attention.0.weight"] = _fan_in_normal(g, (128, D, 1), D, 1.4)
    w[f"{s}.attention.0.bias"] = 0.1 * torch.randn((128,), generator=g)
    bn(f"{s}.attention.2", 128)
    w[f"{s}.attention.3.weight"] = _fan_in_normal(g, (D, 128, 1), 128)
    w[f"{s}.attention.3.bias"] = 0.1 * torch.randn((D,), generator=g)
    out_dim = D * 2 if cfg.resnet_encoder_type == "ASP" else D
    w[f"{s}.fc.weight"] = _fan_in_normal(g, (H, out_dim), out_dim)
    w[f"{s}.fc.bias"] = 0.1 * torch.randn((H,), generator=g)

    # mel decoder (fs2.py:264-278)
    d = "_mel_decoder"
    if cfg.decoder_kind == "fastspeech2":
        w[f"{d}.position_enc"] = _sinusoid_table(cfg.max_mel_len + 1, H).unsqueeze(0)
        fft_stack(d, cfg.dec_layers, scln=cfg.dec_scln)
        w[f"{d}.mel_linear.weight"] = _fan_in_normal(g, (cfg.n_mels, H), H, 2.0)
        w[f"{d}.mel_linear.bias"] = 0.5 * torch.randn((cfg.n_mels,), generator=g)
    elif cfg.decoder_kind == "styletts":
        # StyleTTSDecoder(dim_in=H, style_dim=H, residual_dim=64, dim_out=n_mels) (model.py:238-242; styletts.py:144-179);
        # conv weights stay in weight-norm form (weight_g / weight_v: styletts.py never removes it)
        RD, BN_ = 64, 2 * H

        def wn_conv(prefix, cout, cin, k, bias=True, gain=1.0):
            w[f"{prefix}.weight_v"] = _fan_in_normal(g, (cout, cin, k), cin * k, 1.0)
            w[f"{prefix}.weight_g"] = gain * (0.8 + 0.4 * torch.rand((cout, 1, 1), generator=g))
            if bias:
                w[f"{prefix}.bias"] = 0.1 * torch.randn((cout,), generator=g)

        def inorm(prefix, c):
            w[f"{prefix}.weight"] = 1.0 + 0.1 * torch.randn((c,), generator=g)
            w[f"{prefix}.bias"] = 0.1 * torch.randn((c,), generator=g)

        for i, (ci, co) in enumerate(((H, BN_), (BN_, BN_))):           # encode: ResBlk1d(normalize=True)
            p = f"{d}.encode.{i}"
            wn_conv(f"{p}.conv1", ci, ci, 3, gain=1.4)
            wn_conv(f"{p}.conv2", co, ci, 3, gain=1.4)
            inorm(f"{p}.norm1", ci)
            inorm(f"{p}.norm2", ci)
            if ci != co:
                wn_conv(f"{p}.conv1x1", co, ci, 1, bias=False)
        dims = ((BN_ + RD, BN_), (BN_ + RD, BN_), (BN_ + RD, H), (H, H), (H, H))
        for i, (ci, co) in enumerate(dims):                              # decode: AdainResBlk1d
            p = f"{d}.decode.{i}"
            wn_conv(f"{p}.conv1", co, ci, 3, gain=1.4)
            wn_conv(f"{p}.conv2", co, co, 3, gain=1.4)
            for nm, c in (("norm1", ci), ("norm2", co)):
                w[f"{p}.{nm}.fc.weight"] = 0.5 * torch.randn((2 * c, H), generator=g)   # |style| = 1 -> O(0.5) gamma / beta
                w[f"{p}.{nm}.fc.bias"] = 0.1 * torch.randn((2 * c,), generator=g)
            if ci != co:
                wn_conv(f"{p}.conv1x1", co, ci, 1, bias=False)
        wn_conv(f"{d}.asr_res.0", RD, H, 1)
        inorm(f"{d}.asr_res.1", RD)
        wn_conv(f"{d}.to_out.0", cfg.n_mels, H, 1, gain=2.0)
    else:
        raise NotImplementedError("oracle weights for decoder_kind=%r" % cfg.decoder_kind)

    # vocoder, post-remove_weight_norm form (hifigan.py:93-110, 132-139)
    w.update({f"_meldec.{k}": v for k, v in make_hifigan_weights(cfg.hifigan, g).items()})
    return w


def make_hifigan_weights(h: HifiGanConfig, g: torch.Generator) -> dict:
    w = {}
    C0 = h.upsample_initial_channel
    w["conv_pre.weight"] = _fan_in_normal(g, (C0, 80, 7), 80 * 7)
    w["conv_pre.bias"] = 0.1 * torch.randn((C0,), generator=g)
    ch = C0
    for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
        cin, cout = C0 // (2 ** i), C0 // (2 ** (i + 1))
        # ConvTranspose1d weight is [C_in, C_out, k]; k/u taps reach each output
        w[f"ups.{i}.weight"] = _fan_in_normal(g, (cin, cout, k), cin * (k // u), 1.4)
        w[f"ups.{i}.bias"] = 0.1 * torch.randn((cout,), generator=g)
        ch = cout
        for j, (rk, rd) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
            p = f"resblocks.{i * len(h.resblock_kernel_sizes) + j}"
            names = (["convs1", "convs2"] if h.resblock == "1" else ["convs"])
            for nm in names:
                for di in range(len(rd)):
                    w[f"{p}.{nm}.{di}.weight"] = _fan_in_normal(g, (ch, ch, rk), ch * rk, 0.8)
                    w[f"{p}.{nm}.{di}.bias"] = 0.05 * torch.randn((ch,), generator=g)
    w["conv_post.weight"] = _fan_in_normal(g, (1, ch, 7), ch * 7, 0.7)
    w["conv_post.bias"] = 0.05 * torch.randn((1,), generator=g)
    return w


def make_inputs(cfg: ZeroVoxConfig, B: int, T: int, T_ref: int, seed: int = 7,
                ragged: bool = False, dur_lo: int = 2, dur_hi: int = 10) -> dict:
    """Synthetic batch: phoneme~U{1..27}, puncts~U{0..9}, ref_mel~N(0,1),
    forced durations~U{dur_lo..dur_hi}; ``ragged`` pads a random tail with phoneme_mask."""
    g = torch.Generator().manual_seed(seed)
    x = {
        "phoneme": torch.randint(1, cfg.num_phones, (B, T), generator=g, dtype=torch.int32),
        "puncts": torch.randint(0, cfg.num_puncts, (B, T), generator=g, dtype=torch.int32),
        "ref_mel": torch.randn((B, T_ref, cfg.n_mels), generator=g),
        "duration": torch.randint(dur_lo, dur_hi + 1, (B, T), generator=g, dtype=torch.int32),
    }
    if ragged:
        lens = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
        lens[0] = T
        mask = torch.arange(T)[None, :] >= lens[:, None]
        x["phoneme_mask"] = mask
        x["phoneme"] = x["phoneme"].masked_fill(mask, 0)
        x["puncts"] = x["puncts"].masked_fill(mask, 0)
        x["duration"] = x["duration"].masked_fill(mask, 0)
    return x




def make_speech_like(n: int, seed: int = 1, sr: int = 22050, lead: float = 0.12, tail: float = 0.10) -> np.ndarray:
    """A seeded speech-shaped prompt for the mel front-end: a gliding harmonic stack under a
    syllable-rate envelope plus a -50 dB noise floor, with near-silent (-80 dB) lead-in / tail so that
    `librosa.effects.trim(top_db=40)` has something to cut.  float32 in (-1, 1); pure arithmetic, float64 inside."""
    g = torch.Generator().manual_seed(seed)
    t = np.arange(n, dtype=np.float64) / sr
    f0 = 110.0 + 40.0 * np.sin(2 * np.pi * 0.7 * t + 0.3 * seed)
    phase = 2 * np.pi * np.cumsum(f0) / sr
    amps = torch.rand(24, generator=g, dtype=torch.float64).numpy()
    y = np.zeros(n, dtype=np.float64)
    for h in range(24):
        y += amps[h] / (1.0 + 0.35 * h) * np.sin((h + 1) * phase + h)
    env = 0.55 + 0.45 * np.sin(2 * np.pi * 3.1 * t + seed)
    y = 0.25 * y / np.abs(y).max() * env
    y += 10 ** (-50 / 20) * torch.randn(n, generator=g, dtype=torch.float64).numpy()
    n0, n1 = int(lead * n), int(tail * n)
    gate = np.ones(n)
    gate[:n0] = 1e-4
    gate[n - n1:] = 1e-4
    return (y * gate).astype(np.float32)
