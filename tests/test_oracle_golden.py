"""CPU: the oracle restatement against the committed golden fixtures (which oracle/make_goldens.py produced by
running the reference's own modules — see that script) plus the invariants the reference asserts."""
import dataclasses
import os

import numpy as np
import pytest
import torch

from oracle import zerovox_oracle as zo

CASES = {
    "tiny_forced": zo.ZeroVoxConfig.tiny,
    "tiny_predicted": zo.ZeroVoxConfig.tiny,
    "tiny_longform": zo.ZeroVoxConfig.tiny,
    "medium_forced": zo.ZeroVoxConfig,
    "medium_predicted": zo.ZeroVoxConfig,
    "tiny_styledec": lambda: dataclasses.replace(zo.ZeroVoxConfig.tiny(), decoder_kind="styletts"),
    "medium_styledec": lambda: dataclasses.replace(zo.ZeroVoxConfig(), decoder_kind="styletts"),
}


def load_case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = CASES[name]()
    w = zo.make_weights(cfg, seed=int(g["seed_w"]), dur_bias=float(g["dur_bias"]))
    x = zo.make_inputs(cfg, int(g["B"]), int(g["T"]), int(g["T_ref"]), seed=int(g["seed_x"]),
                       ragged=bool(g["ragged"]), dur_lo=int(g["dur_lo"]), dur_hi=int(g["dur_hi"]))
    return cfg, w, x, g


@pytest.mark.parametrize("name", list(CASES))
def test_forward_matches_reference_golden(golden_dir, name):
    cfg, w, x, g = load_case(golden_dir, name)
    with torch.no_grad():
        wav, mel, mel_len, logd, st = zo.zerovox_forward(cfg, w, dict(x), force_duration=bool(g["force"]))
    # integer outputs: exact
    assert np.array_equal(mel_len.numpy(), g["mel_len"])
    assert np.array_equal(st["_src_index"], g["src_index"])
    assert np.array_equal(st["_pitch_bucket"].numpy(), g["pitch_bucket"])
    assert np.array_equal(st["_energy_bucket"].numpy(), g["energy_bucket"])
    # float outputs: same machine class -> tight tolerance (MKL/oneDNN reduction order may differ across hosts)
    for key, val, tol in (("style_embed", st["style_embed"], 1e-5), ("pitch", st["pitch"], 1e-4),
                          ("energy", st["energy"], 1e-4), ("log_duration", logd, 1e-4), ("mel", mel, 1e-3),
                          ("wav", wav, 1e-3)):
        np.testing.assert_allclose(val.numpy(), g[key], atol=tol, rtol=0, err_msg=key)
    # invariants the reference asserts / guarantees
    if bool(g["force"]):  # utils/export_hifigan.py:125-128
        assert np.array_equal(mel_len.numpy(), x["duration"].clamp(min=0).sum(1).numpy())
    assert wav.shape[1] == mel.shape[2] * cfg.hop_length
    assert float(wav.abs().max()) <= 1.0
    np.testing.assert_allclose(st["style_embed"].norm(dim=-1).numpy(), 1.0, atol=1e-5)  # ResNetSE34V2.py:207-208


@pytest.mark.parametrize("name", ["tiny_forced", "medium_predicted", "tiny_styledec"])
def test_inference_ex_matches_reference_golden(golden_dir, name):
    cfg, w, x, g = load_case(golden_dir, name)
    x1 = {k: v[:1] for k, v in x.items() if k != "phoneme_mask"}
    style = torch.from_numpy(g["style_embed"][:1])
    with torch.no_grad():
        wav, mel_len, logd, mel, mml = zo.zerovox_inference_ex(cfg, w, x1, style, force_duration=bool(g["force"]),
                                                               min_mel_len=int(g["ix_min_mel_len"]))
    assert mel_len == int(g["ix_mel_len"])
    assert wav.shape[0] == mel_len * cfg.hop_length  # model.py:347
    np.testing.assert_allclose(wav.numpy(), g["ix_wav"], atol=1e-3, rtol=0)
    np.testing.assert_allclose(mel.numpy(), g["ix_mel"], atol=1e-3, rtol=0)


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
def test_hifigan_variants_match_reference_golden(golden_dir, v):
    g = np.load(os.path.join(golden_dir, f"hifigan_{v}.npz"))
    h = getattr(zo.HifiGanConfig, v)()
    gen = torch.Generator().manual_seed(int(g["seed"]))
    hw = zo.make_hifigan_weights(h, gen)
    with torch.no_grad():
        wav = zo.hifigan_generator(h, hw, torch.from_numpy(g["mel"]), prefix="")
    np.testing.assert_allclose(wav.numpy(), g["wav"], atol=1e-4, rtol=0)


def test_length_regulator_indices_edge_cases():
    # zero / negative durations, empty utterance, explicit max_len (fs2.py:447-455, 403-423)
    dur = np.array([[2, 0, 3, -1], [0, 0, 0, 0], [1, 1, 1, 1]], dtype=np.int32)
    idx, mel_len = zo.length_regulator_indices(dur)
    assert mel_len.tolist() == [5, 0, 4]
    assert idx.tolist() == [[0, 0, 2, 2, 2], [-1] * 5, [0, 1, 2, 3, -1]]
    idx2, _ = zo.length_regulator_indices(dur, max_len=7)
    assert idx2.shape == (3, 7) and idx2[0].tolist() == [0, 0, 2, 2, 2, -1, -1]


def test_sinusoid_table_matches_formula_rows():
    t = zo.get_sinusoid_encoding_table(5, 8)
    assert t.shape == (5, 8) and torch.all(t[0, 0::2] == 0) and torch.all(t[0, 1::2] == 1)
    np.testing.assert_allclose(t[3, 0].item(), np.sin(3.0), rtol=1e-6)
    np.testing.assert_allclose(t[3, 3].item(), np.cos(3.0 / 10000 ** (2 / 8)), rtol=1e-6)
