"""CPU oracle for the ZeroVOX eval-mode phoneme -> mel -> waveform forward path.

TEST INFRASTRUCTURE ONLY.  This file is the *checker*, never the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  Nothing under ``zerovox_b200/`` does.

It is a plain functional restatement (torch CPU fp32 ops + numpy integer code,
no ``nn.Module``) of the reference's algorithm, keyed by the reference's own
``state_dict`` names.  Every function cites the reference file:line it follows
(paths relative to the upstream repo gooofy/zerovox @ 56a4316).

Pinning status: the reference ships NO tests / golden vectors,
so this restatement is pinned against *outputs of the reference's own modules
run in the build container* — ``oracle/make_goldens.py`` imports
``zerovox.tts.{fs2,hifigan,ResNetSE34V2,model}`` from /root/reference, loads
the same seeded weights, asserts module-vs-restatement agreement and writes
``tests/golden/*.npz`` (committed).  ``tests/test_oracle_golden.py`` re-checks
the restatement against those fixtures on any machine.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # hifigan.py:15


# configuration + seeded synthetic weights / inputs live in the (CUDA-free) zerovox_b200.synthetic module so that
# bench.py's product arm can build its workload without importing the oracle.
from zerovox_b200.synthetic import (HifiGanConfig, ZeroVoxConfig, make_hifigan_weights, make_inputs,  # noqa: E402,F401
                                    make_weights)


def get_sinusoid_encoding_table(n_position: int, d_hid: int) -> torch.Tensor:
    """fs2.py:17-37 — float64 numpy table, cast to fp32 (vectorised, same arithmetic)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    hid = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (hid // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.from_numpy(table.astype(np.float32))



# ----------------------------------------------------------------------------
# FastSpeech2 pieces
# ----------------------------------------------------------------------------
def scln(x, s, affine_w, eps=1e-8):
    """fs2.py:76-90 — unbiased std, (sigma + eps), bias rows first / gain rows second."""
    mu = torch.mean(x, dim=-1, keepdim=True)
    sigma = torch.std(x, dim=-1, keepdim=True)
    y = (x - mu) / (sigma + eps)
    H = x.shape[-1]
    b, g = torch.split(F.linear(s, affine_w), H, dim=-1)
    return g * y + b


def multi_head_attention(w, p, x, spk, slf_attn_mask, n_head, scln_on):
    """fs2.py:133-164 (+ ScaledDotProductAttention 47-58); dropout is identity in eval."""
    B, L, H = x.shape
    dk = H // n_head
    residual = x
    q = F.linear(x, w[f"{p}.w_qs.weight"], w[f"{p}.w_qs.bias"]).view(B, L, n_head, dk)
    k = F.linear(x, w[f"{p}.w_ks.weight"], w[f"{p}.w_ks.bias"]).view(B, L, n_head, dk)
    v = F.linear(x, w[f"{p}.w_vs.weight"], w[f"{p}.w_vs.bias"]).view(B, L, n_head, dk)
    q = q.permute(2, 0, 1, 3).contiguous().view(-1, L, dk)
    k = k.permute(2, 0, 1, 3).contiguous().view(-1, L, dk)
    v = v.permute(2, 0, 1, 3).contiguous().view(-1, L, dk)
    mask = slf_attn_mask.repeat(n_head, 1, 1)
    attn = torch.bmm(q, k.transpose(1, 2))
    attn = attn / np.power(dk, 0.5)
    attn = attn.masked_fill(mask, -np.inf)
    attn = torch.softmax(attn, dim=2)
    out = torch.bmm(attn, v)
    out = out.view(n_head, B, L, dk).permute(1, 2, 0, 3).contiguous().view(B, L, -1)
    out = F.linear(out, w[f"{p}.fc.weight"], w[f"{p}.fc.bias"])
    if scln_on:
        return scln(out + residual, spk, w[f"{p}.layer_norm.affine_layer.linear.weight"])
    return F.layer_norm(out + residual, (H,), w[f"{p}.layer_norm.weight"], w[f"{p}.layer_norm.bias"], 1e-5)


def positionwise_ffn(w, p, x, spk, kernel_size, scln_on):
    """fs2.py:196-209 — Conv1d k=9 (pad 4) -> ReLU -> Conv1d k=1, residual, (SC)LN."""
    H = x.shape[-1]
    residual = x
    o = x.transpose(1, 2)
    o = F.conv1d(o, w[f"{p}.w_1.weight"], w[f"{p}.w_1.bias"], padding=(kernel_size[0] - 1) // 2)
    o = F.relu(o)
    o = F.conv1d(o, w[f"{p}.w_2.weight"], w[f"{p}.w_2.bias"], padding=(kernel_size[1] - 1) // 2)
    o = o.transpose(1, 2)
    if scln_on:
        return scln(o + residual, spk, w[f"{p}.layer_norm.affine_layer.linear.weight"])
    return F.layer_norm(o + residual, (H,), w[f"{p}.layer_norm.weight"], w[f"{p}.layer_norm.bias"], 1e-5)


def fft_block(w, p, x, spk, mask, slf_attn_mask, n_head, kernel_size, scln_on):
    """fs2.py:221-230 — both sub-layers followed by masked_fill(mask, 0)."""
    o = multi_head_attention(w, f"{p}.slf_attn", x, spk, slf_attn_mask, n_head, scln_on)
    o = o.masked_fill(mask.unsqueeze(-1), 0)
    o = positionwise_ffn(w, f"{p}.pos_ffn", o, spk, kernel_size, scln_on)
    o = o.masked_fill(mask.unsqueeze(-1), 0)
    return o


def _pos_table(w, key, L, H, max_len):
    """fs2.py:287-304 / 383-392 — parameter rows when L <= max, recomputed table otherwise (eval)."""
    if L > max_len:
        return get_sinusoid_encoding_table(L, H)[:L, :].to(w[key].device)
    return w[key][0, :L, :]


def encoder(cfg, w, phoneme, puncts, mask):
    """fs2.py:370-401."""
    e = "_phoneme_encoder._encoder"
    x = F.embedding(phoneme.long(), w[f"{e}.src_word_emb.weight"])
    xp = F.embedding(puncts.long(), w[f"{e}.punct_embed.weight"])
    x = torch.cat((x, xp), 2)
    B, T = phoneme.shape
    slf = mask.unsqueeze(1).expand(-1, T, -1)
    x = x + _pos_table(w, f"{e}.position_enc", T, cfg.hidden, cfg.max_txt_len).unsqueeze(0)
    for i in range(cfg.enc_layers):
        x = fft_block(w, f"{e}.layer_stack.{i}", x, None, mask, slf, cfg.enc_heads,
                      cfg.conv_kernel_size, scln_on=False)
    return x


def variance_predictor(cfg, w, p, x, mask):
    """fs2.py:522-563 — conv k3 -> ReLU -> LN -> conv k3 (padding=1) -> ReLU -> LN -> Linear -> mask0."""
    F_ = cfg.vp_filter_size
    K = cfg.vp_kernel_size
    o = F.conv1d(x.transpose(1, 2), w[f"{p}.conv_layer.conv1d_1.conv.weight"],
                 w[f"{p}.conv_layer.conv1d_1.conv.bias"], padding=(K - 1) // 2).transpose(1, 2)
    o = F.relu(o)
    o = F.layer_norm(o, (F_,), w[f"{p}.conv_layer.layer_norm_1.weight"], w[f"{p}.conv_layer.layer_norm_1.bias"], 1e-5)
    o = F.conv1d(o.transpose(1, 2), w[f"{p}.conv_layer.conv1d_2.conv.weight"],
                 w[f"{p}.conv_layer.conv1d_2.conv.bias"], padding=1).transpose(1, 2)
    o = F.relu(o)
    o = F.layer_norm(o, (F_,), w[f"{p}.conv_layer.layer_norm_2.weight"], w[f"{p}.conv_layer.layer_norm_2.bias"], 1e-5)
    o = F.linear(o, w[f"{p}.linear_layer.weight"], w[f"{p}.linear_layer.bias"]).squeeze(-1)
    return o.masked_fill(mask, 0.0)


def bucketize(cfg, pred):
    """fs2.py:639 / 649 — clamp(round(p * (n_bins-1)).long(), 0, n_bins-1); round = half-to-even."""
    return torch.clamp(torch.round(pred * (cfg.ve_n_bins - 1)).long(), min=0, max=cfg.ve_n_bins - 1)


def duration_round(log_d):
    """fs2.py:678-681."""
    return torch.clamp(torch.round(torch.exp(log_d) - 1), min=0)


def length_regulator_indices(duration: np.ndarray, max_len: int | None = None):
    """fs2.py:432-459 + pad 403-423, as pure integer index arithmetic.

    Returns (src_index int32 [B, L_max] with -1 at zero-padded frames, mel_len int64 [B]).
    Row i is repeated max(int(d_i), 0) times (fs2.py:451-452).
    """
    dur = np.maximum(duration.astype(np.int64), 0)
    mel_len = dur.sum(axis=1)
    L = int(max_len) if max_len else int(mel_len.max()) if mel_len.size else 0
    B, T = dur.shape
    idx = np.full((B, L), -1, dtype=np.int32)
    for b in range(B):
        rep = np.repeat(np.arange(T, dtype=np.int32), dur[b])
        idx[b, : len(rep)] = rep[:L]
    return idx, mel_len.astype(np.int64)


def length_regulate(x, duration, max_len=None):
    idx, mel_len = length_regulator_indices(duration.detach().cpu().numpy(), max_len)
    idx_t = torch.from_numpy(idx).long().to(x.device)
    B, L = idx_t.shape
    g = torch.gather(x, 1, idx_t.clamp(min=0).unsqueeze(-1).expand(B, L, x.shape[-1]))
    g = g.masked_fill((idx_t < 0).unsqueeze(-1), 0.0)
    return g, torch.from_numpy(mel_len).to(x.device), idx


def fs2_encoder(cfg, w, x, style_embed, force_duration=False):
    """FS2Encoder.forward, fs2.py:732-775 + VarianceAdaptor.forward 652-693 (eval)."""
    phoneme, puncts = x["phoneme"], x["puncts"]
    mask = x["phoneme_mask"] if "phoneme_mask" in x else torch.zeros_like(phoneme, dtype=torch.bool)
    feats = encoder(cfg, w, phoneme, puncts, mask)
    feats = feats + style_embed.expand_as(feats)  # all positions, padded too (fs2.py:740-741)
    va = "_phoneme_encoder._variance_adaptor"
    log_d = variance_predictor(cfg, w, f"{va}.duration_predictor", feats, mask)
    pitch = variance_predictor(cfg, w, f"{va}.pitch_predictor", feats, mask)
    pb = bucketize(cfg, pitch)
    feats = feats + F.embedding(pb, w[f"{va}.pitch_embedding.weight"])
    energy = variance_predictor(cfg, w, f"{va}.energy_predictor", feats, mask)
    eb = bucketize(cfg, energy)
    feats = feats + F.embedding(eb, w[f"{va}.energy_embedding.weight"])
    xprime = feats
    if force_duration:
        dur = x["duration"]
        out, mel_len, idx = length_regulate(feats, dur)
        # fs2.py:748, 772: a collated batch carries 'mel_mask' (data.py:85-93) and it becomes the mask of the forced path
        masks = x["mel_mask"].unsqueeze(2).expand(-1, -1, out.shape[2]) if "mel_mask" in x else None
    else:
        dur = duration_round(log_d)
        out, mel_len, idx = length_regulate(feats, dur)
        L = out.shape[1]
        mel_mask = torch.arange(L, device=mel_len.device)[None, :] >= mel_len[:, None]  # fs2.py:565-573
        masks = mel_mask.unsqueeze(2).expand(-1, -1, out.shape[2])
    return {"pitch": pitch, "energy": energy, "log_duration": log_d, "mel_len": mel_len,
            "features": out, "masks": masks,
            # extras for stage-wise parity tests (not in the reference dict)
            "_xprime": xprime, "_pitch_bucket": pb, "_energy_bucket": eb,
            "_duration_rounded": dur, "_src_index": idx}


def fs2_decoder(cfg, w, enc_seq, mask, spk):
    """FS2Decoder.forward, fs2.py:281-315 (eval: no truncation to max_seq_len)."""
    d = "_mel_decoder"
    B, L, H = enc_seq.shape
    slf = mask.unsqueeze(1).expand(-1, L, -1)
    o = enc_seq + _pos_table(w, f"{d}.position_enc", L, H, cfg.max_mel_len).unsqueeze(0)
    for i in range(cfg.dec_layers):
        o = fft_block(w, f"{d}.layer_stack.{i}", o, spk, mask, slf, cfg.dec_heads,
                      cfg.conv_kernel_size, scln_on=cfg.dec_scln)
    return F.linear(o, w[f"{d}.mel_linear.weight"], w[f"{d}.mel_linear.bias"])


# ----------------------------------------------------------------------------
# ResNetSE34V2 speaker embedding
# ----------------------------------------------------------------------------
def _bn(w, p, x):
    return F.batch_norm(x, w[f"{p}.running_mean"], w[f"{p}.running_var"], w[f"{p}.weight"], w[f"{p}.bias"],
                        training=False, eps=1e-5)


def se_basic_block(w, p, x, stride):
    """ResNetSE34V2.py:83-99 — conv->ReLU->BN (!), conv->BN, SE gate, +residual, ReLU."""
    out = F.conv2d(x, w[f"{p}.conv1.weight"], None, stride=stride, padding=1)
    out = F.relu(out)
    out = _bn(w, f"{p}.bn1", out)
    out = F.conv2d(out, w[f"{p}.conv2.weight"], None, padding=1)
    out = _bn(w, f"{p}.bn2", out)
    y = out.mean(dim=(2, 3))  # AdaptiveAvgPool2d(1), ResNetSE34V2.py:63-67
    y = torch.sigmoid(F.linear(F.relu(F.linear(y, w[f"{p}.se.fc.0.weight"], w[f"{p}.se.fc.0.bias"])),
                               w[f"{p}.se.fc.2.weight"], w[f"{p}.se.fc.2.bias"]))
    out = out * y[:, :, None, None]
    if f"{p}.downsample.0.weight" in w:
        res = F.conv2d(x, w[f"{p}.downsample.0.weight"], None, stride=stride)
        res = _bn(w, f"{p}.downsample.1", res)
    else:
        res = x
    return F.relu(out + res)


def speaker_embed(cfg, w, ref_mel):
    """ResNetSE34V2.forward, ResNetSE34V2.py:176-212 (log_input=False, model.py:223)."""
    s = "_spkemb"
    x = ref_mel.transpose(1, 2)
    x = F.instance_norm(x, eps=1e-5).unsqueeze(1)
    x = F.conv2d(x, w[f"{s}.conv1.weight"], w[f"{s}.conv1.bias"], padding=1)
    x = F.relu(x)
    x = _bn(w, f"{s}.bn1", x)
    for li, nblocks in enumerate(cfg.resnet_layers, start=1):
        for bi in range(nblocks):
            stride = 2 if (li > 1 and bi == 0) else 1
            x = se_basic_block(w, f"{s}.layer{li}.{bi}", x, stride)
    x = x.reshape(x.size(0), -1, x.size(-1))
    a = F.conv1d(x, w[f"{s}.attention.0.weight"], w[f"{s}.attention.0.bias"])
    a = F.relu(a)
    a = _bn(w, f"{s}.attention.2", a)
    a = F.conv1d(a, w[f"{s}.attention.3.weight"], w[f"{s}.attention.3.bias"])
    a = torch.softmax(a, dim=2)
    if cfg.resnet_encoder_type == "SAP":
        x = torch.sum(x * a, dim=2)
    else:
        mu = torch.sum(x * a, dim=2)
        sg = torch.sqrt((torch.sum((x ** 2) * a, dim=2) - mu ** 2).clamp(min=1e-5))
        x = torch.cat((mu, sg), 1)
    x = F.linear(x, w[f"{s}.fc.weight"], w[f"{s}.fc.bias"])
    x = F.normalize(x, p=2, dim=1)
    return x.unsqueeze(1)


# ----------------------------------------------------------------------------
# HiFi-GAN generator
# ----------------------------------------------------------------------------
def _get_padding(k, d=1):  # hifigan.py:22-23
    return int((k * d - d) / 2)


def hifigan_generator(h: HifiGanConfig, w: dict, mel: torch.Tensor, prefix: str = "_meldec.") -> torch.Tensor:
    """Generator.forward, hifigan.py:114-130; ResBlock1 49-56, ResBlock2 78-82.  mel [B,80,L] -> [B,1,256L]."""
    g = lambda k: w[prefix + k]
    nk = len(h.resblock_kernel_sizes)
    x = F.conv1d(mel, g("conv_pre.weight"), g("conv_pre.bias"), padding=3)
    for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, g(f"ups.{i}.weight"), g(f"ups.{i}.bias"), stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
            p = f"resblocks.{i * nk + j}"
            r = x
            for di, d in enumerate(rd):
                xt = F.leaky_relu(r, LRELU_SLOPE)
                if h.resblock == "1":
                    xt = F.conv1d(xt, g(f"{p}.convs1.{di}.weight"), g(f"{p}.convs1.{di}.bias"),
                                  dilation=d, padding=_get_padding(rk, d))
                    xt = F.leaky_relu(xt, LRELU_SLOPE)
                    xt = F.conv1d(xt, g(f"{p}.convs2.{di}.weight"), g(f"{p}.convs2.{di}.bias"),
                                  padding=_get_padding(rk, 1))
                else:
                    xt = F.conv1d(xt, g(f"{p}.convs.{di}.weight"), g(f"{p}.convs.{di}.bias"),
                                  dilation=d, padding=_get_padding(rk, d))
                r = xt + r
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)  # default slope 0.01 (hifigan.py:126)
    x = F.conv1d(x, g("conv_post.weight"), g("conv_post.bias"), padding=3)
    return torch.tanh(x)



# ----------------------------------------------------------------------------
# StyleTTS mel decoder (alternative decoder_kind, the shipped default models)
# ----------------------------------------------------------------------------
def _wn(w, p):
    """torch.nn.utils.weight_norm with dim=0: w = g * v / ||v|| per output channel (styletts.py:28-34)."""
    v, g = w[f"{p}.weight_v"], w[f"{p}.weight_g"]
    return g * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)


def _wn_conv(w, p, x, padding):
    return F.conv1d(x, _wn(w, p), w.get(f"{p}.bias"), padding=padding)


def _adain(w, p, x, s):
    """AdaIN1d.forward, styletts.py:87-92."""
    h = F.linear(s, w[f"{p}.fc.weight"], w[f"{p}.fc.bias"]).unsqueeze(-1)
    gamma, beta = torch.chunk(h, 2, dim=1)
    return (1 + gamma) * F.instance_norm(x, eps=1e-5) + beta


def styletts_decoder(cfg, w, enc_seq, mask, spk_emb, prefix="_mel_decoder"):
    """StyleTTSDecoder.forward, styletts.py:181-205 (ResBlk1d 55-69, AdainResBlk1d 124-139); dropout is identity in
    eval, `mask` is ignored by the reference.  enc_seq [B,L,H], spk_emb [B,1,H] -> mel [B,L,n_mels]."""
    d = prefix
    x0 = enc_seq.transpose(1, 2)
    s = spk_emb.squeeze(1)
    x = x0
    for i in range(2):   # encode: ResBlk1d(normalize=True, downsample='none')
        p = f"{d}.encode.{i}"
        sc = _wn_conv(w, f"{p}.conv1x1", x, 0) if f"{p}.conv1x1.weight_v" in w else x
        r = F.instance_norm(x, weight=w[f"{p}.norm1.weight"], bias=w[f"{p}.norm1.bias"], eps=1e-5)
        r = _wn_conv(w, f"{p}.conv1", F.leaky_relu(r, 0.2), 1)
        r = F.instance_norm(r, weight=w[f"{p}.norm2.weight"], bias=w[f"{p}.norm2.bias"], eps=1e-5)
        r = _wn_conv(w, f"{p}.conv2", F.leaky_relu(r, 0.2), 1)
        x = (sc + r) / math.sqrt(2)
    asr = F.instance_norm(_wn_conv(w, f"{d}.asr_res.0", x0, 0), weight=w[f"{d}.asr_res.1.weight"],
                          bias=w[f"{d}.asr_res.1.bias"], eps=1e-5)
    res = True
    for i in range(5):   # decode: AdainResBlk1d; block 2 is the (non-)upsampling one after which the residual stops
        p = f"{d}.decode.{i}"
        if res:
            x = torch.cat([x, asr], dim=1)
        r = _wn_conv(w, f"{p}.conv1", F.leaky_relu(_adain(w, f"{p}.norm1", x, s), 0.2), 1)
        r = _wn_conv(w, f"{p}.conv2", F.leaky_relu(_adain(w, f"{p}.norm2", r, s), 0.2), 1)
        sc = _wn_conv(w, f"{p}.conv1x1", x, 0) if f"{p}.conv1x1.weight_v" in w else x
        x = (r + sc) / math.sqrt(2)
        if i == 2:
            res = False
    x = _wn_conv(w, f"{d}.to_out.0", x, 0)
    return x.transpose(1, 2)


def mel_decoder(cfg, w, features, dec_mask, spk_emb):
    """Dispatch on decoder_kind (model.py:225-244)."""
    if cfg.decoder_kind == "styletts":
        return styletts_decoder(cfg, w, features, dec_mask, spk_emb)
    return fs2_decoder(cfg, w, features, dec_mask, spk_emb)

# ----------------------------------------------------------------------------
# model container
# ----------------------------------------------------------------------------
def zerovox_forward(cfg, w, x, force_duration=False, style_embed=None):
    """ZeroVox.forward eval path, model.py:260-290, with the *intended* HiFi-GAN tail.

    The reference's own eval tail (model.py:298-304) raises with hifigan.Generator
    (ParallelWaveGAN leftovers .mean/.scale/.pqmf/c=); the tail used here is
    ``wav = _meldec(mel.transpose(1,2)).squeeze(1)``, which is what
    utils/export_hifigan.py:109-151 consumes.  Returns (wav [B,256*L], mel [B,80,L],
    mel_len [B] int64, log_duration [B,T]) plus a dict of stage tensors.
    """
    se = speaker_embed(cfg, w, x["ref_mel"]) if style_embed is None else style_embed
    pred = fs2_encoder(cfg, w, x, se, force_duration=force_duration)
    masks = pred["masks"]
    L = pred["features"].shape[1]
    if masks is None:  # model.py:269-273
        dec_mask = ~(torch.arange(L, device=pred["mel_len"].device).expand(len(pred["mel_len"]), L) < pred["mel_len"].unsqueeze(1))
    else:
        dec_mask = masks[:, :, 0]
    mel = mel_decoder(cfg, w, pred["features"], dec_mask, se)
    if masks is not None and mel.size(0) > 1:  # model.py:283-285
        mel = mel.masked_fill(masks[:, :, : mel.shape[-1]], 0)
    mel_t = mel.transpose(1, 2)
    wav = hifigan_generator(cfg.hifigan, w, mel_t).squeeze(1)
    stages = dict(pred)
    stages["style_embed"] = se
    stages["dec_mask"] = dec_mask
    return wav, mel_t, pred["mel_len"], pred["log_duration"], stages


def zerovox_inference_ex(cfg, w, x, style_embed, force_duration=False, min_mel_len=689):
    """ZeroVox.inference_ex, model.py:308-347 (batch = 1).  Returns the reference 4-tuple and the
    updated ``_min_mel_len`` (the reference mutates self._min_mel_len, model.py:331-335)."""
    pred = fs2_encoder(cfg, w, x, style_embed, force_duration=force_duration)
    L = pred["features"].shape[1]
    dec_mask = ~(torch.arange(L, device=pred["mel_len"].device).expand(1, L) < pred["mel_len"].unsqueeze(1))
    mel = mel_decoder(cfg, w, pred["features"], dec_mask, style_embed)
    mel_len = int(pred["mel_len"][0])
    mel = mel[0]
    if mel_len < min_mel_len:
        mel = F.pad(mel, (0, 0, 0, min_mel_len - mel_len))
    elif mel_len > min_mel_len:
        min_mel_len = mel_len
    wav = hifigan_generator(cfg.hifigan, w, mel.T.unsqueeze(0))[0, 0]
    mel = mel.transpose(0, 1)
    return wav[: mel_len * cfg.hop_length], mel_len, pred["log_duration"], mel[:, :mel_len], min_mel_len
