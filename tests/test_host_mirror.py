"""CPU: the host-side mirror keeps the reference's constructor signature, attribute names and state_dict keys,
and refuses (loudly) to run anywhere but on the CUDA engine."""
import inspect
import os

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


def test_checkpoint_ingestion_like_synthesize(tmp_path, monkeypatch):
    """The on-disk formats ZeroVoxTTS.__init__ / get_meldec read (synthesize.py:78-88, model.py:86-118): a Lightning
    .ckpt ({'hyper_parameters', 'state_dict'}) loaded with strict=False + extra kwargs, and a vocoder directory with
    config.json + generator.ckpt['generator'] in weight-norm form under $CACHED_PATH_ZEROVOX/model_repo/<name>/."""
    import json
    from zerovox_b200.tts import Generator, AttrDict, get_meldec
    cfg = zo.ZeroVoxConfig.tiny()
    w = zo.make_weights(cfg, seed=0)
    # vocoder repo in the HF-cache layout of model.py:66-82
    name = "zerovox-hifigan-test"
    repo = tmp_path / "model_repo" / name
    repo.mkdir(parents=True)
    (repo / "config.json").write_text(json.dumps(cfg.hifigan.as_json_dict()))
    raw_gen = Generator(AttrDict(cfg.hifigan.as_json_dict()))          # still weight-normed, like upstream checkpoints
    torch.save({"generator": raw_gen.state_dict()}, repo / "generator.ckpt")
    monkeypatch.setenv("CACHED_PATH_ZEROVOX", str(tmp_path))
    gen = get_meldec(name)
    assert "conv_pre.weight" in gen.state_dict() and not gen.training
    torch.testing.assert_close(gen.state_dict()["conv_pre.weight"], raw_gen._engine_state_dict()["conv_pre.weight"])
    assert os.path.isdir(repo) and get_meldec(str(repo)).state_dict().keys() == gen.state_dict().keys()   # directory form
    # Lightning checkpoint: hyper-parameters minus the ignored ones, acoustic-model weights only (no _meldec.* keys)
    kw = zerovox_kwargs(cfg)
    hp = {k: v for k, v in kw.items() if k not in ("meldec_model",)}
    sd = {k: v for k, v in w.items() if not k.startswith("_meldec.")}
    ckpt = tmp_path / "epoch=0001.ckpt"
    torch.save({"hyper_parameters": hp, "state_dict": sd}, ckpt)
    zv = ZeroVox.load_from_checkpoint(lang="en", meldec_model=name, sampling_rate=cfg.sampling_rate, hop_length=cfg.hop_length,
                                      checkpoint_path=str(ckpt), infer_device="cpu", map_location="cpu", strict=False,
                                      verbose=False, betas=(0.0, 0.99), eps=1e-9)
    assert isinstance(zv._meldec, Generator)
    for k, v in sd.items():
        torch.testing.assert_close(zv.state_dict()[k], v)
    with pytest.raises(FileNotFoundError):
        get_meldec("no-such-model")
