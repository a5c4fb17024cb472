"""``zerovox_b200.patch()`` — accelerate the REFERENCE's own classes in place (north_star: "demo.py and
train_tts.py call it unchanged").

The reference's ``zerovox.tts.model.ZeroVox`` keeps every parameter, its constructor, ``state_dict``, optimiser hooks and
its training-mode ``forward`` (model.py:260-293: returns ``pred`` for the loss).  Only two methods are rebound, and only for
the case the CUDA engine covers:

    eval mode  AND  parameters on a CUDA device  AND  ``_meldec`` is a ``hifigan.Generator``
        ZeroVox.forward       (model.py:260-306)  -> zvx_spkemb / zvx_encode / zvx_length_regulate / zvx_decode / zvx_vocode
        ZeroVox.inference_ex  (model.py:308-347)  -> the same stages, batch-1 semantics incl. the stateful ``_min_mel_len``

Everything else — ``self.training``, CPU parameters, no vocoder — falls through to the reference's own code, untouched
(including its own failure modes: the upstream eval tail raises with ``hifigan.Generator``, model.py:298-304).

The engine receives the module's ``state_dict`` under the reference's keys; the copy is refreshed whenever a parameter's
version counter changed (an optimiser step, ``load_state_dict``) or the module moved to another device, so
train -> eval -> train cycles of utils/train_tts.py see current weights.
"""
from __future__ import annotations

import importlib

import torch
import torch.nn as nn

from .engine import Engine, EngineConfig

_ORIG = {}


def config_from_reference(zv, tensor_core_policy: int = 1) -> EngineConfig:
    """Engine hyper-parameters read off a reference ``ZeroVox`` instance (module structure and parameter shapes; the
    Lightning ``hparams`` are not needed)."""
    enc = zv._phoneme_encoder._encoder
    va = zv._phoneme_encoder._variance_adaptor
    cfg = EngineConfig(tensor_core_policy=tensor_core_policy)
    cfg.num_phones = enc.src_word_emb.weight.shape[0] - 1
    cfg.emb_dim = enc.src_word_emb.weight.shape[1]
    cfg.num_puncts = enc.punct_embed.weight.shape[0] - 1
    cfg.punct_emb_dim = enc.punct_embed.weight.shape[1]
    cfg.max_txt_len = enc.position_enc.shape[1] - 1
    cfg.enc_layers = len(enc.layer_stack)
    cfg.enc_heads = enc.layer_stack[0].slf_attn.n_head
    w1 = enc.layer_stack[0].pos_ffn.w_1
    w2 = enc.layer_stack[0].pos_ffn.w_2
    cfg.conv_filter_size = w1.weight.shape[0]
    cfg.conv_kernel_size = (w1.weight.shape[2], w2.weight.shape[2])
    c1 = va.duration_predictor.conv_layer.conv1d_1.conv
    cfg.vp_filter_size, cfg.vp_kernel_size = c1.weight.shape[0], c1.weight.shape[2]
    cfg.ve_n_bins = va.pitch_embedding.weight.shape[0]
    dec = zv._mel_decoder
    if hasattr(dec, "layer_stack"):
        cfg.decoder_kind = "fastspeech2"
        cfg.max_mel_len = dec.position_enc.shape[1] - 1
        cfg.dec_layers = len(dec.layer_stack)
        cfg.dec_heads = dec.layer_stack[0].slf_attn.n_head
        cfg.dec_scln = bool(dec.layer_stack[0].slf_attn.scln)
        cfg.n_mels = dec.mel_linear.weight.shape[0]
    else:
        cfg.decoder_kind = "styletts"
        cfg.n_mels = zv._spkemb.n_mels
    spk = zv._spkemb
    cfg.resnet_layers = tuple(len(getattr(spk, f"layer{i}")) for i in (1, 2, 3, 4))
    cfg.resnet_num_filters = tuple(getattr(spk, f"layer{i}")[0].conv1.weight.shape[0] for i in (1, 2, 3, 4))
    cfg.resnet_encoder_type = spk.encoder_type
    if zv._meldec is not None and hasattr(zv._meldec, "h"):
        cfg.set_hifigan(zv._meldec.h)
    return cfg


def _engine_state_dict(zv) -> dict:
    """The module's state_dict with the vocoder convs in plain-``weight`` form (weight norm folded when still attached,
    hifigan.py:132-139)."""
    sd = {k: v for k, v in zv.state_dict().items() if not k.startswith("_meldec.")}
    if zv._meldec is not None:
        for name, m in zv._meldec.named_modules():
            if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d)):
                w = torch._weight_norm(m.weight_v, m.weight_g, 0) if hasattr(m, "weight_g") else m.weight
                sd["_meldec." + name + ".weight"] = w.detach()
                sd["_meldec." + name + ".bias"] = m.bias.detach()
    return sd


class _Accel:
    """Engine handle + weight freshness for one reference ``ZeroVox`` instance."""

    def __init__(self):
        self.engine = None
        self.stamp = None

    def get(self, zv, device) -> Engine:
        stamp = (str(device), sum(int(t._version) for t in zv.state_dict(keep_vars=True).values()),
                 sum(1 for _ in zv.parameters()))
        if self.engine is None or self.engine.device != device:
            self.engine = Engine(config_from_reference(zv, getattr(zv, "_zvx_tensor_core_policy", 1)), device)
            self.stamp = None
        if stamp != self.stamp:
            self.engine.load_weights(_engine_state_dict(zv))
            self.stamp = stamp
        return self.engine


def _engine_for(zv):
    """The engine when the accelerated case applies to this call, else None (-> the reference's own code)."""
    if zv.training or getattr(zv, "_meldec", None) is None or not hasattr(zv._meldec, "h"):
        return None
    p = next(zv.parameters(), None)
    if p is None or p.device.type != "cuda":
        return None
    dev = p.device if p.device.index is not None else torch.device("cuda", torch.cuda.current_device())
    acc = zv.__dict__.get("_zvx_accel")
    if acc is None:
        acc = zv.__dict__["_zvx_accel"] = _Accel()
    return acc.get(zv, dev)


def patch(reference_model_module=None):
    """Rebind the eval-mode CUDA ``forward`` / ``inference_ex`` of the reference's ``ZeroVox``.  ``reference_model_module``:
    the imported ``zerovox.tts.model`` (default: imported by name).  Idempotent; ``unpatch()`` restores the originals."""
    from .tts.model import engine_forward, engine_inference_ex
    mod = reference_model_module or importlib.import_module("zerovox.tts.model")
    cls = mod.ZeroVox
    if cls in _ORIG:
        return cls
    orig_forward, orig_inference_ex = cls.forward, cls.inference_ex
    _ORIG[cls] = (orig_forward, orig_inference_ex)

    def forward(self, x, force_duration=False, normalize_before=True, **kw):
        eng = _engine_for(self)
        if eng is None:
            return orig_forward(self, x, force_duration=force_duration, normalize_before=normalize_before)
        return engine_forward(eng, x, force_duration=force_duration, **kw)

    def inference_ex(self, x, style_embed, normalize_before=True, force_duration=False, **kw):
        eng = _engine_for(self)
        if eng is None:
            return orig_inference_ex(self, x, style_embed, normalize_before=normalize_before, force_duration=force_duration)
        return engine_inference_ex(self, eng, x, style_embed, force_duration=force_duration, **kw)

    forward.__doc__ = (orig_forward.__doc__ or "") + "\n[zerovox_b200.patch: eval-mode CUDA calls run in the B200 engine]"
    cls.forward, cls.inference_ex = forward, inference_ex
    return cls


def unpatch():
    for cls, (f, ix) in list(_ORIG.items()):
        cls.forward, cls.inference_ex = f, ix
        del _ORIG[cls]
