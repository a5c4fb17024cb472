"""ZeroVox model container with the reference's public interface (zerovox/tts/model.py:86-118, 158-351):
same constructor kwargs, ``forward`` / ``inference_ex`` / ``inference`` signatures and return tuples, same
attribute names (``_phoneme_encoder``, ``_spkemb``, ``_mel_decoder``, ``_meldec``, ``_min_mel_len``,
``_hop_length``, ``hparams``) and the same ``state_dict`` keys, so zerovox/tts/synthesize.py, zerovox/demo.py and
utils/export_hifigan.py drive it unchanged.  Eval-mode arithmetic runs in the CUDA engine; there is no CPU or
training path here (see INTEGRATION.md for how training keeps using the reference modules).
"""
from __future__ import annotations

import json
import os
import time
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.nn as nn

from ._context import EngineContext
from .fs2 import FS2Decoder, FS2Encoder
from .hifigan import Generator
from .styletts import StyleTTSDecoder
from .ResNetSE34V2 import ResNetSE34V2
from .symbols import Symbols

DEFAULT_MELDEC_MODEL_NAME = "zerovox-hifigan-vctk-v2-en-1"  # model.py:84


class AttrDict(dict):
    """dict with attribute access, for HiFi-GAN config.json (model.py:39-42)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def model_cache_path(model: str, relpath: str) -> Path:
    """Cache layout of model.py:66-82: $CACHED_PATH_ZEROVOX (default ~/.cache/zerovox)/model_repo/<model>/<relpath>.
    Files must already be there — this engine never downloads."""
    cache = Path(os.getenv("CACHED_PATH_ZEROVOX", Path.home() / ".cache" / "zerovox"))
    path = cache / "model_repo" / model / relpath
    if not path.exists():
        raise FileNotFoundError(f"{path} not found (zerovox_b200 does not download models; place the file there)")
    return path


def get_meldec(modelspec, infer_device="cpu", verbose=False):
    """model.py:86-118: build the Generator from <dir>/config.json + generator.ckpt['generator'] (weight-norm form),
    eval(), remove_weight_norm()."""
    if os.path.isdir(modelspec):
        config_path, gen_path = Path(modelspec) / "config.json", Path(modelspec) / "generator.ckpt"
    else:
        config_path = model_cache_path(str(modelspec), "config.json")
        gen_path = model_cache_path(str(modelspec), "generator.ckpt")
    if verbose:
        print("meldec: using config    : ", config_path)
        print("meldec: using checkpoint: ", gen_path)
    with open(config_path) as f:
        config = AttrDict(json.loads(f.read()))
    device = torch.device(infer_device)
    generator = Generator(config).to(device)
    state = torch.load(gen_path, map_location=device)
    generator.load_state_dict(state["generator"])
    generator.eval()
    generator.remove_weight_norm()
    return generator.to(device)


_CHILD_ROLES = {"_phoneme_encoder": "encoder", "_mel_decoder": "decoder", "_spkemb": "spkemb", "_meldec": "vocoder"}


class ZeroVox(nn.Module):

    def __init__(self, symbols: Symbols, meldec_model, sampling_rate, hop_length, n_mels, lr, weight_decay,
                 max_epochs, warmup_epochs, betas, eps, embed_dim, punct_embed_dim, dpe_embed_dim, emb_reduction,
                 max_mel_len, max_txt_len, fs2enc_layer, fs2enc_head, fs2enc_dropout, vp_filter_size,
                 vp_kernel_size, vp_dropout, ve_n_bins, resnet_layers, resnet_num_filters, resnet_encoder_type,
                 decoder_kind, decoder_n_layers, decoder_n_head, decoder_conv_filter_size, decoder_conv_kernel_size,
                 decoder_dropout, decoder_scln, verbose=False):
        super().__init__()
        hp = dict(locals())
        for k in ("self", "__class__", "meldec_model", "verbose"):
            hp.pop(k, None)
        self.hparams = SimpleNamespace(**hp)  # save_hyperparameters(ignore=[...]) stand-in (model.py:204)
        object.__setattr__(self, "_shared_ctx", EngineContext())

        emb_size = embed_dim + punct_embed_dim
        self._phoneme_encoder = FS2Encoder(
            symbols=symbols, max_txt_len=max_txt_len, embed_dim=embed_dim, encoder_layer=fs2enc_layer,
            encoder_head=fs2enc_head, conv_filter_size=decoder_conv_filter_size,
            conv_kernel_size=decoder_conv_kernel_size, encoder_dropout=fs2enc_dropout,
            punct_embed_dim=punct_embed_dim, vp_filter_size=vp_filter_size, vp_kernel_size=vp_kernel_size,
            vp_dropout=vp_dropout, ve_n_bins=ve_n_bins)
        self._spkemb = ResNetSE34V2(layers=resnet_layers, num_filters=resnet_num_filters, nOut=emb_size,
                                    encoder_type=resnet_encoder_type, n_mels=n_mels, log_input=False)
        if decoder_kind == "fastspeech2":
            self._mel_decoder = FS2Decoder(
                dec_max_seq_len=max_mel_len, dec_hidden=emb_size, dec_n_layers=decoder_n_layers,
                dec_n_head=decoder_n_head, dec_conv_filter_size=decoder_conv_filter_size,
                dec_conv_kernel_size=decoder_conv_kernel_size, dec_dropout=decoder_dropout, dec_scln=decoder_scln,
                n_mel_channels=n_mels, spk_emb_size=emb_size)
        elif decoder_kind == "styletts":
            self._mel_decoder = StyleTTSDecoder(dim_in=emb_size, style_dim=emb_size, residual_dim=64, dim_out=n_mels)
        else:
            raise Exception(f"unknown decoder kind: '{decoder_kind}'")
        self._meldec = get_meldec(modelspec=meldec_model, verbose=verbose) if meldec_model else None
        self._min_mel_len = 689  # model.py:254
        self._hop_length = hop_length
        self._verbose = verbose

    # children share one engine handle ------------------------------------------------------------------
    def __setattr__(self, name, value):
        super().__setattr__(name, value)
        role = _CHILD_ROLES.get(name)
        if role is not None and isinstance(value, nn.Module) and hasattr(value, "_fill_config"):
            self._shared_ctx.attach(role, value)

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._shared_ctx.mark_stale()
        return r

    def load_state_dict(self, state_dict, strict=True, assign=False):
        r = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._shared_ctx.mark_stale()
        return r

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **kwargs):
        """Lightning-style .ckpt ingestion (synthesize.py:78-88): ``hyper_parameters`` + ``state_dict``;
        kwargs override / complete the hyper-parameters, unknown ones are dropped."""
        import inspect
        ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(kwargs)
        allowed = set(inspect.signature(cls.__init__).parameters) - {"self"}
        model = cls(**{k: v for k, v in hp.items() if k in allowed})
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        return model

    # forward -------------------------------------------------------------------------------------------
    def forward(self, x, force_duration=False, normalize_before=True, *, pad_to=None, zero_padded_mel=None,
                vocoder_groups=None, on_group=None):
        """Batched eval forward (model.py:260-306); see :func:`engine_forward` for the return tuple and the keyword-only
        extensions."""
        if self.training:
            raise NotImplementedError("ZeroVox.forward in training mode is outside the zerovox_b200 hot path; "
                                      "zerovox_b200.patch() keeps training on the reference modules")
        if self._meldec is None:
            raise RuntimeError("ZeroVox.forward: no vocoder (_meldec is None)")
        eng = self._shared_ctx.get(next(self.parameters()).device)
        return engine_forward(eng, x, force_duration=force_duration, pad_to=pad_to, zero_padded_mel=zero_padded_mel,
                              vocoder_groups=vocoder_groups, on_group=on_group)

    def inference_ex(self, x, style_embed, normalize_before=True, force_duration=False, *, vocoder_chunk_frames=None):
        """Batch-1 path (model.py:308-347); see :func:`engine_inference_ex`."""
        eng = self._shared_ctx.get(next(self.parameters()).device)
        return engine_inference_ex(self, eng, x, style_embed, force_duration=force_duration,
                                   vocoder_chunk_frames=vocoder_chunk_frames)

    def inference(self, x, style_embed, normalize_before=True):
        wav, mel_len, log_duration, _ = self.inference_ex(x=x, style_embed=style_embed,
                                                           normalize_before=normalize_before)
        return wav, mel_len, log_duration


# ---------------------------------------------------------------------------------------------------------------------
# the two eval-mode call sequences over the C ABI, shared by the mirror class above and by zerovox_b200.patch() (which
# rebinds them on the reference's own ZeroVox class)
# ---------------------------------------------------------------------------------------------------------------------
def group_bounds(n: int, groups: int) -> list[tuple[int, int]]:
    """Consecutive utterance groups (the vocoder's delivery units): [(first, end), ...].  groups > 0: that many nearly equal
    groups; groups < 0: |groups| groups of halving size (n/2, n/4, ..., the last two equal) — the transfer of a group hides
    behind the vocoding of ALL later groups, so only the small last group's transfer is exposed."""
    g = int(groups or 1)
    if g >= 0:
        g = max(1, min(g, n))
        return [(n * i // g, n * (i + 1) // g) for i in range(g)]
    g = max(1, min(-g, n))
    cuts, left = [0], n
    for i in range(g - 1):
        take = max(1, left // 2) if left > (g - 1 - i) else 1
        cuts.append(cuts[-1] + take)
        left -= take
    cuts.append(n)
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


def host_delivery(wav_host: torch.Tensor, stream: "torch.cuda.Stream | None" = None):
    """``on_group`` callback for ``forward(..., vocoder_groups=G)``: copies every finished utterance group of the padded
    waveform batch to ``wav_host`` (page-locked, [B, >= L*hop]) on a side stream, so that the device-to-host transfer of group i
    runs while group i + 1 is vocoded; with groups of halving size only the last, small group's copy is exposed.  The caller
    synchronises (``callback.stream.synchronize()`` or a device synchronise) before reading ``wav_host``."""
    side = stream

    def on_group(i, g0, g1, wav, mel, mel_len):
        nonlocal side
        if side is None:
            side = torch.cuda.Stream(device=wav.device)
            on_group.stream = side
        side.wait_stream(torch.cuda.current_stream(wav.device))
        with torch.cuda.stream(side):
            wav_host[g0:g1, : wav.shape[1]].copy_(wav[g0:g1], non_blocking=True)
        wav.record_stream(side)

    on_group.stream = side
    return on_group


def engine_forward(eng, x, force_duration=False, *, pad_to=None, zero_padded_mel=None, vocoder_groups=None, on_group=None):
    """Batched eval forward (model.py:260-306).  Returns (wav [B, L_max*hop], mel [B, n_mels, L_max], mel_len int64 [B],
    log_duration [B, T]) — the tuple utils/export_hifigan.py:109-151 consumes.  The reference's own eval tail
    (model.py:298-304) is ParallelWaveGAN leftover code that raises with hifigan.Generator; the intended semantics
    ``wav = _meldec(mel.transpose(1,2)).squeeze(1)`` are built.

    Keyword-only extensions used by zerovox_b200.parallel so that a shard reproduces the unsharded batch: ``pad_to`` — an
    int, or a callable ``(local_L_max, mel_len_host) -> L`` — is the frame count to pad the batch to (>= its own maximum);
    ``zero_padded_mel`` overrides the reference's batch-size dependent zero-fill of padded mel frames (model.py:283-285
    applies it iff a mel mask exists and B > 1 — B being the GLOBAL batch there).  ``vocoder_groups`` = G > 1 vocodes the
    batch in G consecutive utterance groups (the generator treats utterances independently: same samples) and calls
    ``on_group(i, first, end, wav, mel, mel_len)`` after enqueuing each, so that a caller can ship finished waveforms (NCCL
    gather, device-to-host copy) while the next group is still being computed."""
    dev = eng.device
    mask = x["phoneme_mask"].to(dev, non_blocking=True) if "phoneme_mask" in x else None
    forced = x["duration"].to(dev, non_blocking=True) if force_duration else None
    # model.py:263-265 as one call: the speaker net runs on a side stream next to the encoder's FFT blocks
    style, r = eng.spkemb_encode(x["ref_mel"].to(dev, non_blocking=True), x["phoneme"].to(dev, non_blocking=True),
                                 x["puncts"].to(dev, non_blocking=True), mask, forced)
    L = r["L_max"]
    if pad_to is not None:
        L = max(L, int(pad_to(L, r["mel_len_host"]) if callable(pad_to) else pad_to))
    feats = eng.length_regulate(r["xprime"], r["duration_rounded"], L)
    # fs2.py:748, 772 + model.py:264-285: the mel mask is x['mel_mask'] when the collated batch carries one (the
    # utils/export_hifigan.py flow, forced durations), else the one derived from predicted durations; with forced
    # durations and no 'mel_mask' there is none.  The mel is zero-filled at padded frames iff a mask exists and B > 1.
    dec_mask = None
    if force_duration and "mel_mask" in x:
        dec_mask = x["mel_mask"].to(dev, non_blocking=True)
        if dec_mask.shape[1] != L:
            raise RuntimeError(f"x['mel_mask'] covers {dec_mask.shape[1]} frames, the durations give {L} "
                               "(the reference fails on this batch too: fs2.py:772, model.py:279)")
    zero_pad = zero_padded_mel
    if zero_pad is None:
        zero_pad = ((not force_duration) or "mel_mask" in x) and feats.shape[0] > 1
    _, mel = eng.decode(feats, style, mask=dec_mask, mel_len=r["mel_len"], zero_padded_mel=bool(zero_pad), want_blc=False)
    B = mel.shape[0]
    if vocoder_groups and abs(int(vocoder_groups)) > 1 and B > 1:
        wav = torch.empty((B, L * eng.cfg.hop_length), device=dev, dtype=torch.float32)
        for i, (g0, g1) in enumerate(group_bounds(B, vocoder_groups)):
            eng.vocode(mel[g0:g1], out=wav[g0:g1])
            if on_group is not None:
                on_group(i, g0, g1, wav, mel, r["mel_len"])
    else:
        wav = eng.vocode(mel).squeeze(1)
        if on_group is not None:
            on_group(0, 0, B, wav, mel, r["mel_len"])
    return wav, mel, r["mel_len"], r["log_duration"]


def engine_inference_ex(owner, eng, x, style_embed, force_duration=False, *, vocoder_chunk_frames=None):
    """Batch-1 path (model.py:308-347): zero-pads the mel to the stateful ``owner._min_mel_len`` before vocoding and trims
    the waveform to mel_len*hop.  Returns (wav, mel_len, log_duration, mel [n_mels, mel_len]).
    ``vocoder_chunk_frames`` (keyword-only extension, long-form inputs): vocode in chunks of that many mel frames with a
    14-frame discarded halo — same waveform, bounded workspace."""
    start_time = time.time()
    dev = eng.device
    forced = x["duration"].to(dev) if force_duration else None
    mask = x["phoneme_mask"].to(dev) if "phoneme_mask" in x else None
    r = eng.encode(x["phoneme"].to(dev), x["puncts"].to(dev), style_embed.to(dev), mask, forced)
    if len(r["mel_len_host"]) != 1:
        raise RuntimeError("inference_ex is the batch-1 path (model.py:325); use forward() for batches")
    mel_len = int(r["mel_len_host"][0])
    feats = eng.length_regulate(r["xprime"], r["duration_rounded"], r["L_max"])
    pe_time = time.time()
    _, mel = eng.decode(feats, style_embed.to(dev), mel_len=r["mel_len"], want_blc=False)
    dec_time = time.time()
    if mel_len < owner._min_mel_len:
        padded = torch.zeros((1, mel.shape[1], owner._min_mel_len), device=dev, dtype=mel.dtype)
        padded[:, :, :mel_len] = mel
    else:
        owner._min_mel_len = max(owner._min_mel_len, mel_len)
        padded = mel
    if vocoder_chunk_frames:
        wav = eng.vocode_chunked(padded, int(vocoder_chunk_frames), 14)[0, 0]
    else:
        wav = eng.vocode(padded)[0, 0]
    if getattr(owner, "_verbose", False):
        torch.cuda.synchronize(dev)
        now = time.time()
        print(f"synthesis timing stats: pe={pe_time - start_time}s, dec={dec_time - pe_time}s, "
              f"meldec={now - dec_time}s")
    return wav[: mel_len * owner._hop_length], mel_len, r["log_duration"], mel[0, :, :mel_len]
