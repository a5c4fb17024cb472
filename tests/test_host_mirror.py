"""CPU: the host-side mirror keeps the reference's constructor signature, attribute names and state_dict keys,
and refuses (loudly) to run anywhere but on the CUDA engine."""
import inspect

import pytest
import torch

from oracle import zerovox_oracle as zo
from zerovox_b200.testing import build_generator, build_model, zerovox_kwargs
from zerovox_b200.tts import ZeroVox
from zerovox_b200.tts.symbols import Symbols

# ZeroVox.__init__ parameters of the reference (zerovox/tts/model.py:159-201)
REFERENCE_CTOR = ["symbols", "meldec_model", "sampling_rate", "hop_length", "n_mels", "lr", "weight_decay",
                  "max_epochs", "warmup_epochs", "betas", "eps", "embed_dim", "punct_embed_dim", "dpe_embed_dim",
                  "emb_reduction", "max_mel_len", "max_txt_len", "fs2enc_layer", "fs2enc_head", "fs2enc_dropout",
                  "vp_filter_size", "vp_kernel_size", "vp_dropout", "ve_n_bins", "resnet_layers",
                  "resnet_num_filters", "resnet_encoder_type", "decoder_kind", "decoder_n_layers", "decoder_n_head",
                  "decoder_conv_filter_size", "decoder_conv_kernel_size", "decoder_dropout", "decoder_scln", "verbose"]


def test_constructor_and_method_signatures_match_reference():
    assert list(inspect.signature(ZeroVox.__init__).parameters)[1:] == REFERENCE_CTOR
    fwd = inspect.signature(ZeroVox.forward).parameters
    positional = [n for n, p in fwd.items() if p.kind is not inspect.Parameter.KEYWORD_ONLY]
    assert positional == ["self", "x", "force_duration", "normalize_before"]   # extensions are keyword-only
    iex = inspect.signature(ZeroVox.inference_ex).parameters
    assert [n for n, p in iex.items() if p.kind is not inspect.Parameter.KEYWORD_ONLY] == [
        "self", "x", "style_embed", "normalize_before", "force_duration"]
    assert list(inspect.signature(ZeroVox.inference).parameters) == ["self", "x", "style_embed", "normalize_before"]


def _styledec_tiny():
    import dataclasses
    return dataclasses.replace(zo.ZeroVoxConfig.tiny(), decoder_kind="styletts")


@pytest.mark.parametrize("cfgf", [zo.ZeroVoxConfig.tiny, zo.ZeroVoxConfig, _styledec_tiny])
def test_state_dict_keys_match_reference(cfgf):
    cfg = cfgf()
    w = zo.make_weights(cfg, seed=0)  # keyed like the reference state_dict (validated by oracle/make_goldens.py)
    zv = build_model(cfg, w)          # strict on everything but the unused torchfb buffers
    keys = set(zv.state_dict().keys())
    extra = {k for k in keys - set(w) if "torchfb" not in k}
    assert not extra, sorted(extra)[:5]
    assert not (set(w) - keys), sorted(set(w) - keys)[:5]
    assert {"_spkemb.torchfb.0.flipped_filter", "_spkemb.torchfb.1.spectrogram.window",
            "_spkemb.torchfb.1.mel_scale.fb"} <= keys
    for attr in ("_phoneme_encoder", "_spkemb", "_mel_decoder", "_meldec", "_min_mel_len", "_hop_length", "hparams"):
        assert hasattr(zv, attr)
    assert zv._min_mel_len == 689 and zv.hparams.embed_dim == cfg.emb_dim
    for k, v in w.items():
        assert zv.state_dict()[k].shape == v.shape, k


def test_generator_accepts_weight_norm_checkpoints():
    h = zo.HifiGanConfig.v2()
    from zerovox_b200.tts import Generator, AttrDict
    gen = Generator(AttrDict(h.as_json_dict()))
    raw = gen.state_dict()
    assert "conv_pre.weight_g" in raw and "ups.0.weight_v" in raw and "resblocks.0.convs1.0.weight_g" in raw
    folded = gen._engine_state_dict()
    v, g = raw["ups.0.weight_v"], raw["ups.0.weight_g"]
    ref = v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))  # w = g * v / ||v||, dim 0 = in-channel for ConvTranspose1d
    torch.testing.assert_close(folded["ups.0.weight"], ref)
    gen.remove_weight_norm()
    assert "conv_pre.weight" in gen.state_dict() and "conv_pre.weight_g" not in gen.state_dict()
    torch.testing.assert_close(gen.state_dict()["ups.0.weight"], ref)


def test_no_cpu_fallback_and_no_training_path():
    cfg = zo.ZeroVoxConfig.tiny()
    zv = build_model(cfg, zo.make_weights(cfg, seed=0))
    x = zo.make_inputs(cfg, 1, 5, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        zv(x, force_duration=True)            # CPU module -> loud failure, never a silent CPU path
    with pytest.raises(RuntimeError, match="CUDA"):
        zv._spkemb(x["ref_mel"])
    zv.train()
    with pytest.raises(NotImplementedError):
        zv(x)


def test_unknown_decoder_kind_raises_like_reference():
    cfg = zo.ZeroVoxConfig.tiny()
    kw = zerovox_kwargs(cfg)
    kw["decoder_kind"] = "bogus"
    with pytest.raises(Exception, match="unknown decoder kind"):
        ZeroVox(**kw)


def test_symbols_table():
    s = Symbols("'-abc", " ,.")
    assert s.num_phones == 5 and s.num_puncts == 4
    assert s.encode_phone("'") == 0 and s.decode_phone(2) == "a"
    assert s.encode_punct(Symbols.NO_PUNCT) == 0 and s.encode_punct(",") == 2 and s.is_punct(".") and not s.is_phone(".")
