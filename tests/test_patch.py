"""zerovox_b200.patch(): the reference's OWN ZeroVox class keeps training / CPU on its own code and only its eval-mode CUDA
forward / inference_ex are rebound (north_star "demo.py and train_tts.py call it unchanged").

CPU part (runs wherever the unmodified reference modules are importable: /root/reference or the oracle/_ref copy): the
training-mode forward still returns the reference's `pred` dict (model.py:290-293), the constructor takes the kwargs of
utils/train_tts.py:202-241, state_dict round-trips, eval-on-CPU falls through to the reference code, and the engine
configuration is read off the module correctly.  GPU part: the patched reference class gives the engine's results.
"""
import dataclasses

import numpy as np
import pytest
import torch

from oracle import reference_modules as rm
from oracle import zerovox_oracle as zo

needs_ref = pytest.mark.skipif(not rm.available(), reason="reference modules not available (oracle/build_ref.py not run)")


def train_batch(cfg, B=2, T=6, seed=5):
    g = torch.Generator().manual_seed(seed)
    x = zo.make_inputs(cfg, B, T, 16, seed=seed, ragged=True, dur_lo=1, dur_hi=4)
    mel_len = x["duration"].clamp(min=0).sum(1).to(torch.int32)
    L = int(mel_len.max())
    x["mel_len"] = mel_len
    x["mel_mask"] = torch.arange(L)[None, :] >= mel_len[:, None]
    x["pitch"] = torch.rand(B, T, generator=g)
    x["energy"] = torch.rand(B, T, generator=g)
    return x


@needs_ref
def test_patch_keeps_training_and_cpu_on_the_reference():
    import zerovox_b200
    model_mod, _, _ = rm.import_reference()
    orig_forward = model_mod.ZeroVox.forward
    cls = zerovox_b200.patch(model_mod)
    try:
        assert cls is model_mod.ZeroVox and cls.forward is not orig_forward
        assert zerovox_b200.patch(model_mod) is cls                       # idempotent
        cfg = zo.ZeroVoxConfig.tiny()
        w = zo.make_weights(cfg, seed=1)
        zv = rm.build_reference_model(cfg, w)                              # kwargs of utils/train_tts.py:202-241
        # state_dict round trip through the patched class
        sd = zv.state_dict()
        zv2 = rm.build_reference_model(cfg, zo.make_weights(cfg, seed=2))
        zv2.load_state_dict(sd)
        for k, v in zv2.state_dict().items():
            assert torch.equal(v, sd[k]), k
        # training mode: the reference's own forward, returning `pred` (model.py:290-293) with gradients
        zv.train()
        x = train_batch(cfg)
        torch.manual_seed(0)
        pred = zv(x)
        assert isinstance(pred, dict) and {"mel", "pitch", "energy", "log_duration", "mel_len", "features", "masks"} <= set(pred)
        assert pred["mel"].requires_grad and pred["mel"].shape[0] == 2 and pred["mel"].shape[2] == cfg.n_mels
        torch.manual_seed(0)
        ref_pred = orig_forward(zv, x)
        assert torch.equal(pred["mel"], ref_pred["mel"])                  # same code path, same RNG stream
        pred["mel"].sum().backward()
        assert zv._mel_decoder.mel_linear.weight.grad is not None
        # eval on CPU: not the accelerated case -> the reference's own eval tail, which fails the way upstream does
        zv.eval()
        with pytest.raises(AttributeError), torch.no_grad():
            zv({k: v for k, v in x.items()}, force_duration=True)
        # engine configuration read off the module
        from zerovox_b200.patching import config_from_reference
        ec = config_from_reference(zv)
        assert (ec.emb_dim, ec.punct_emb_dim, ec.max_txt_len, ec.max_mel_len) == (cfg.emb_dim, cfg.punct_emb_dim, cfg.max_txt_len, cfg.max_mel_len)
        assert (ec.enc_layers, ec.enc_heads, ec.dec_layers, ec.dec_heads) == (cfg.enc_layers, cfg.enc_heads, cfg.dec_layers, cfg.dec_heads)
        assert (ec.vp_filter_size, ec.vp_kernel_size, ec.ve_n_bins) == (cfg.vp_filter_size, cfg.vp_kernel_size, cfg.ve_n_bins)
        assert ec.conv_filter_size == cfg.conv_filter_size and tuple(ec.conv_kernel_size) == tuple(cfg.conv_kernel_size)
        assert tuple(ec.resnet_layers) == tuple(cfg.resnet_layers) and tuple(ec.resnet_num_filters) == tuple(cfg.resnet_num_filters)
        assert ec.num_phones == cfg.num_phones and ec.num_puncts == cfg.num_puncts and ec.dec_scln == cfg.dec_scln
        assert ec.hg_upsample_initial_channel == cfg.hifigan.upsample_initial_channel and ec.hop_length == cfg.hop_length
        assert ec.decoder_kind == "fastspeech2" and ec.resnet_encoder_type == cfg.resnet_encoder_type
        st = rm.build_reference_model(dataclasses.replace(cfg, decoder_kind="styletts"), zo.make_weights(dataclasses.replace(cfg, decoder_kind="styletts"), seed=4))
        assert config_from_reference(st).decoder_kind == "styletts"
        ec.to_c()                                                           # valid for the C ABI
    finally:
        zerovox_b200.unpatch()
    assert model_mod.ZeroVox.forward is orig_forward


@needs_ref
@pytest.mark.gpu
def test_patched_reference_runs_the_engine_on_cuda():
    import zerovox_b200
    from zerovox_b200.testing import build_model
    model_mod, _, _ = rm.import_reference()
    zerovox_b200.patch(model_mod)
    try:
        cfg = zo.ZeroVoxConfig.tiny()
        w = zo.make_weights(cfg, seed=1, dur_bias=float(np.log(4.0)))
        x = zo.make_inputs(cfg, 3, 11, 24, seed=7, ragged=True, dur_lo=0, dur_hi=5)
        zv = rm.build_reference_model(cfg, w).to("cuda:0").eval()
        mirror = build_model(cfg, w, device="cuda:0")
        with torch.no_grad():
            out = zv({k: v.to("cuda:0") for k, v in x.items()}, force_duration=True)
            ref = mirror(dict(x), force_duration=True)
            owav, omel, olen, ologd, _ = zo.zerovox_forward(cfg, w, dict(x), force_duration=True)
        for a, b in zip(out, ref):
            assert torch.equal(a, b)                                       # the same engine calls
        assert torch.equal(out[2].cpu(), olen) and float((out[0].cpu() - owav).abs().max()) < 2e-2
        # a weight update (optimiser step in train mode) must reach the engine on the next eval call
        zv.train()
        with torch.no_grad():
            zv._mel_decoder.mel_linear.bias.add_(1.0)
        zv.eval()
        with torch.no_grad():
            out2 = zv({k: v.to("cuda:0") for k, v in x.items()}, force_duration=True)
        assert float((out2[1] - out[1] - 1.0).abs().max()) < 1e-3          # mel shifted by the bias change
        # batch-1 inference_ex with the reference's stateful _min_mel_len
        x1 = {k: v[:1].to("cuda:0") for k, v in x.items() if k != "phoneme_mask"}
        zv._min_mel_len = 40
        with torch.no_grad():
            style = zv._spkemb(x1["ref_mel"])
            wav, n, logd, mel = zv.inference_ex(x1, style_embed=style, force_duration=True)
        assert wav.shape[0] == n * cfg.hop_length and mel.shape == (cfg.n_mels, n)
    finally:
        zerovox_b200.unpatch()
