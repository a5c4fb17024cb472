"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes) and through the reference-shaped
mirror modules, against the CPU oracle on the same seeded inputs and against the committed golden fixtures.

Tolerances (stated per stage, fp32 unless marked):
  * integer outputs (mel_len, forced durations, LengthRegulator indices and gathered rows): bit-exact;
  * fp32 FMA path (tensor_core_policy=0; encoder + variance predictors always): |err| <= 2e-4 * max|ref| + 2e-5;
  * TF32 tensor-core path (policy 1; decoder / vocoder / speaker net; operands rounded to nearest TF32, fp32
    accumulation): mel |err| <= 3.5e-3 * max|ref|, wav |err| <= 1.8e-2, rel-RMS <= 3e-3 — at most 2x the worst values
    measured on B200 over this file's cases (mel 1.9e-3 * max, wav 9.4e-3: the tiny random-weight models are the worst; at
    the benchmarked sizes, tests/test_gpu_fullsize.py, mel 1.75e-3 * max / wav 7.4e-3 / rel-RMS 9.5e-4);
  * pitch / energy buckets and predicted durations: exact wherever the oracle's float input to the rounding step is
    more than 1e-3 away from a rounding boundary (reported otherwise).
  * end-to-end runs under policy 1: the speaker net runs in TF32, so its style vector differs from the reference's by
    ~1e-3; the encoder adds that vector to every phoneme before the variance predictors, whose outputs are ROUNDED to
    buckets / durations (fs2.py:639, 649, 678-681) — a 1e-3 nudge can flip one bucket and change a whole phoneme's
    frames.  That is a property of the reference's arithmetic, not a kernel error, so the policy-1 end-to-end tests
    (a) check the engine's style vector against the reference's within the TF32 tolerance and (b) compare everything
    downstream with the oracle fed that same style vector (stage-wise parity, SURVEY.md §7 "hard parts").
"""
import dataclasses
import os

import numpy as np
import pytest
import torch

from oracle import zerovox_oracle as zo
from zerovox_b200.testing import build_generator, build_model

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.numel() == 0:
        return 0.0, 0.0
    return (a - b).abs().max().item(), b.abs().max().item()


def check(name, got, ref, rtol, atol=2e-5, rms=None):
    err, mag = rel_err(got, ref)
    a, b = torch.as_tensor(got).double().cpu(), torch.as_tensor(ref).double().cpu()
    rel_rms = float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp(min=1e-30)) if a.numel() else 0.0
    ok = err <= rtol * mag + atol and (rms is None or rel_rms <= rms)
    print(f"  {name:28s} max|diff|={err:.3e} max|ref|={mag:.3e} tol={rtol * mag + atol:.3e} rel-RMS={rel_rms:.2e} "
          f"{'ok' if ok else 'FAIL'}")
    assert ok, f"{name}: {err:.3e} > {rtol * mag + atol:.3e} (rel-RMS {rel_rms:.2e}, bar {rms})"


FP32 = dict(rtol=2e-4)
TC_MEL = dict(rtol=3.5e-3, rms=3e-3)
TC_WAV = dict(rtol=0.0, atol=1.8e-2, rms=3e-3)


def tol(policy, kind):
    if policy == 0:
        return FP32
    return TC_WAV if kind == "wav" else TC_MEL


class Case:
    def __init__(self, cfg, seed_w=0, dur_bias=None):
        self.cfg = cfg
        self.w = zo.make_weights(cfg, seed=seed_w, dur_bias=dur_bias)
        self.models = {}

    def model(self, policy):
        if policy not in self.models:
            self.models[policy] = build_model(self.cfg, self.w, device=DEV, tensor_core_policy=policy)
        return self.models[policy]


@pytest.fixture(scope="module")
def tiny():
    return Case(zo.ZeroVoxConfig.tiny(), seed_w=1, dur_bias=float(np.log(4.0)))


@pytest.fixture(scope="module")
def medium():
    return Case(zo.ZeroVoxConfig(), seed_w=0, dur_bias=float(np.log(4.0)))


def to_dev(x):
    return {k: v.to(DEV) for k, v in x.items()}


# ---------------------------------------------------------------------------------------------- stages
@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("which,B,T_ref", [("tiny", 3, 24), ("tiny", 1, 9), ("medium", 2, 48), ("medium", 3, 131)])
def test_speaker_embedding(tiny, medium, which, B, T_ref, policy):
    case = tiny if which == "tiny" else medium
    x = zo.make_inputs(case.cfg, B, 4, T_ref, seed=3)
    with torch.no_grad():
        ref = zo.speaker_embed(case.cfg, case.w, x["ref_mel"])
        got = case.model(policy)._spkemb(x["ref_mel"].to(DEV))
    assert got.shape == ref.shape == (B, 1, case.cfg.hidden)
    check("style_embed", got, ref, **tol(policy, "mel"))
    np.testing.assert_allclose(got.norm(dim=-1).cpu().numpy(), 1.0, atol=1e-5)


@pytest.mark.parametrize("which,B,T,T_ref,forced", [("tiny", 3, 11, 24, False), ("medium", 4, 33, 131, True), ("medium", 32, 128, 440, True)])
def test_spkemb_encode_equals_the_two_calls(tiny, medium, which, B, T, T_ref, forced):
    """zvx_spkemb_encode (speaker net on the engine's side stream next to the encoder's FFT blocks, joined where the style
    vector enters) gives what zvx_spkemb followed by zvx_encode gives, bit for bit — twice in a row (the side stream, its
    events and the speaker net's own workspace are reused)."""
    case = tiny if which == "tiny" else medium
    x = to_dev(zo.make_inputs(case.cfg, B, T, T_ref, seed=13, ragged=True))
    eng = case.model(1)._shared_ctx.get(torch.device(DEV))
    fd = x["duration"] if forced else None
    style = eng.spkemb(x["ref_mel"])
    r = eng.encode(x["phoneme"], x["puncts"], style, x["phoneme_mask"], fd)
    for _ in range(2):
        style2, r2 = eng.spkemb_encode(x["ref_mel"], x["phoneme"], x["puncts"], x["phoneme_mask"], fd)
        assert torch.equal(style2, style)
        for k in ("pitch", "energy", "log_duration", "duration_rounded", "mel_len", "xprime"):
            assert torch.equal(r2[k], r[k]), k
        assert r2["L_max"] == r["L_max"] and r2["mel_len_host"] == r["mel_len_host"]


@pytest.mark.parametrize("which,B,T,ragged", [("tiny", 3, 11, True), ("tiny", 1, 1, False), ("tiny", 2, 30, False),
                                              ("medium", 2, 12, True), ("medium", 4, 33, True)])
def test_encoder_and_variance_adaptor(tiny, medium, which, B, T, ragged):
    """Encoder + variance predictors always run in fp32 FMA (both policies share this path)."""
    case = tiny if which == "tiny" else medium
    cfg = case.cfg
    x = zo.make_inputs(cfg, B, T, 16, seed=5, ragged=ragged)
    style = torch.nn.functional.normalize(torch.randn(B, 1, cfg.hidden, generator=torch.Generator().manual_seed(1)), dim=-1)
    with torch.no_grad():
        ref = zo.fs2_encoder(cfg, case.w, dict(x), style, force_duration=False)
    eng = case.model(1)._shared_ctx.get(torch.device(DEV))
    r = eng.encode(x["phoneme"].to(DEV), x["puncts"].to(DEV), style.to(DEV),
                   x["phoneme_mask"].to(DEV) if ragged else None, None)
    check("log_duration", r["log_duration"], ref["log_duration"], **FP32)
    check("pitch", r["pitch"], ref["pitch"], **FP32)
    # buckets / durations: exact away from rounding boundaries
    def boundary_safe(v):
        return (v - torch.floor(v) - 0.5).abs() > 1e-3
    pb = zo.bucketize(cfg, r["pitch"].cpu())
    safe = boundary_safe(ref["pitch"] * (cfg.ve_n_bins - 1))
    assert torch.equal(pb[safe], ref["_pitch_bucket"][safe])
    if torch.equal(pb, ref["_pitch_bucket"]):
        check("energy", r["energy"], ref["energy"], **FP32)
        eb = zo.bucketize(cfg, r["energy"].cpu())
        safe_e = boundary_safe(ref["energy"] * (cfg.ve_n_bins - 1))
        assert torch.equal(eb[safe_e], ref["_energy_bucket"][safe_e])
        if torch.equal(eb, ref["_energy_bucket"]):
            check("xprime", r["xprime"], ref["_xprime"], **FP32)
    else:
        print("  pitch bucket flipped at a rounding boundary; energy comparison skipped")
    dref = ref["_duration_rounded"]
    safe_d = boundary_safe(torch.exp(ref["log_duration"]) - 1)
    assert torch.equal(r["duration_rounded"].cpu().float()[safe_d], dref[safe_d])
    if torch.equal(r["duration_rounded"].cpu().float(), dref):
        assert r["mel_len"].cpu().tolist() == ref["mel_len"].tolist() == r["mel_len_host"]
        assert r["L_max"] == int(ref["mel_len"].max())


def test_forced_duration_passthrough_and_lengths(tiny):
    cfg = tiny.cfg
    x = zo.make_inputs(cfg, 3, 11, 16, seed=9, ragged=True, dur_lo=0, dur_hi=5)
    style = torch.zeros(3, 1, cfg.hidden)
    style[:, :, 0] = 1.0
    eng = tiny.model(1)._shared_ctx.get(torch.device(DEV))
    r = eng.encode(x["phoneme"].to(DEV), x["puncts"].to(DEV), style.to(DEV), x["phoneme_mask"].to(DEV),
                   x["duration"].to(DEV))
    assert torch.equal(r["duration_rounded"].cpu(), x["duration"])
    assert r["mel_len"].cpu().tolist() == x["duration"].clamp(min=0).sum(1).tolist()  # export_hifigan.py:125-128


@pytest.mark.parametrize("B,T,H,seed", [(1, 1, 96, 0), (3, 17, 96, 1), (5, 128, 528, 2), (2, 700, 528, 3)])
def test_length_regulator_bit_exact(tiny, medium, B, T, H, seed):
    case = tiny if H == 96 else medium
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, H, generator=g)
    dur = torch.randint(-1, 9, (B, T), generator=g, dtype=torch.int32)  # includes 0 and negative (-> 0)
    if B > 1:
        dur[1] = 0  # an empty utterance
    ref_idx, ref_len = zo.length_regulator_indices(dur.numpy())
    L = ref_idx.shape[1]
    eng = case.model(1)._shared_ctx.get(torch.device(DEV))
    for L_max in {L, L + 5}:  # exact fit and explicit max_len padding (fs2.py:440-443)
        feats, idx = eng.length_regulate(x.to(DEV), dur.to(DEV), L_max, want_index=True)
        ridx, _ = zo.length_regulator_indices(dur.numpy(), max_len=L_max)
        assert np.array_equal(idx.cpu().numpy(), ridx)
        ref_feats, _, _ = zo.length_regulate(x, dur, max_len=L_max)
        assert torch.equal(feats.cpu(), ref_feats)  # pure row copy: bit-exact


@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("which,B,L", [("tiny", 3, 35), ("tiny", 1, 93), ("medium", 2, 59), ("medium", 3, 130)])
def test_decoder(tiny, medium, which, B, L, policy):
    case = tiny if which == "tiny" else medium
    cfg = case.cfg
    g = torch.Generator().manual_seed(4)
    mel_len = torch.randint(max(1, L // 2), L + 1, (B,), generator=g)
    mel_len[0] = L
    mask = torch.arange(L)[None, :] >= mel_len[:, None]
    feats = torch.randn(B, L, cfg.hidden, generator=g).masked_fill(mask.unsqueeze(-1), 0.0)
    style = torch.nn.functional.normalize(torch.randn(B, 1, cfg.hidden, generator=g), dim=-1)
    with torch.no_grad():
        ref = zo.fs2_decoder(cfg, case.w, feats, mask, style)
        mel, m2 = case.model(policy)._mel_decoder(feats.to(DEV), mask.to(DEV), style.to(DEV))
    check("mel", mel, ref, **tol(policy, "mel"))
    eng = case.model(policy)._shared_ctx.get(torch.device(DEV))
    blc, bcl = eng.decode(feats.to(DEV), style.to(DEV), mel_len=mel_len.to(DEV), zero_padded_mel=True)
    refz = ref.masked_fill(mask.unsqueeze(-1), 0.0)
    check("mel (zero-padded, BLC)", blc, refz, **tol(policy, "mel"))
    assert torch.equal(bcl, blc.transpose(1, 2))


@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("which,B,L", [("tiny", 3, 35), ("tiny", 1, 7), ("medium", 2, 59), ("medium", 3, 130)])
def test_styletts_decoder(which, B, L, policy):
    """StyleTTSDecoder.forward (styletts.py:181-205) stand-alone: the mirror module with reference-keyed weight-norm
    parameters against the oracle restatement (pinned to the reference module by the *_styledec goldens)."""
    from zerovox_b200.tts.styletts import StyleTTSDecoder
    cfg = golden_cfg(which + "_st")
    w = zo.make_weights(cfg, seed=3)
    g = torch.Generator().manual_seed(17)
    feats = torch.randn(B, L, cfg.hidden, generator=g)
    style = torch.nn.functional.normalize(torch.randn(B, 1, cfg.hidden, generator=g), dim=-1)
    dec = StyleTTSDecoder(dim_in=cfg.hidden, style_dim=cfg.hidden, residual_dim=64, dim_out=cfg.n_mels)
    dec.load_state_dict({k[len("_mel_decoder."):]: v for k, v in w.items() if k.startswith("_mel_decoder.")})
    dec._ctx.tensor_core_policy = policy
    dec = dec.eval().to(DEV)
    with torch.no_grad():
        ref = zo.styletts_decoder(cfg, w, feats, None, style)
        got, none = dec(feats.to(DEV), None, style.to(DEV))
    assert none is None and got.shape == ref.shape == (B, L, cfg.n_mels)
    check("styletts mel", got, ref, **tol(policy, "mel"))


@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("v,B,L", [("v2", 2, 9), ("v1", 1, 7), ("v3", 2, 5), ("v2", 3, 70), ("v2", 1, 1),
                                   ("v1", 2, 40), ("v3", 3, 33), ("v2", 1, 300)])
def test_vocoder_variants(golden_dir, v, B, L, policy):
    h = getattr(zo.HifiGanConfig, v)()
    g = torch.Generator().manual_seed(11)
    hw = zo.make_hifigan_weights(h, g)
    mel = torch.randn((B, 80, L), generator=g)
    gen = build_generator(h, {"_meldec." + k: t for k, t in hw.items()}).to(DEV)
    gen._ctx.tensor_core_policy = policy
    with torch.no_grad():
        ref = zo.hifigan_generator(h, hw, mel, prefix="")
        got = gen(mel.to(DEV))
    assert got.shape == ref.shape == (B, 1, L * 256)
    check(f"wav {v}", got, ref, **tol(policy, "wav"))
    if (B, L) == (2, 9) or (v, B, L) == ("v1", 1, 7):
        gold = np.load(os.path.join(golden_dir, f"hifigan_{v}.npz"))
        if B == 2 and L == 9:  # the fixture produced by hifigan.Generator itself
            check(f"wav {v} vs reference golden", got, gold["wav"], **tol(policy, "wav"))
    # unbatched [80, L] input like inference_ex (model.py:337)
    with torch.no_grad():
        one = gen(mel[0].to(DEV))
    assert one.shape == (1, L * 256)
    assert torch.equal(one, got[0])


# ---------------------------------------------------------------------------------------------- end to end
GOLDEN_CASES = {"tiny_forced": "tiny", "tiny_predicted": "tiny", "tiny_longform": "tiny", "medium_forced": "medium",
                "medium_predicted": "medium", "tiny_styledec": "tiny_st", "medium_styledec": "medium_st"}


def golden_cfg(kind):
    cfg = zo.ZeroVoxConfig.tiny() if kind.startswith("tiny") else zo.ZeroVoxConfig()
    return dataclasses.replace(cfg, decoder_kind="styletts") if kind.endswith("_st") else cfg


@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_forward_against_reference_goldens(golden_dir, name, policy):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = golden_cfg(GOLDEN_CASES[name])
    w = zo.make_weights(cfg, seed=int(g["seed_w"]), dur_bias=float(g["dur_bias"]))
    x = zo.make_inputs(cfg, int(g["B"]), int(g["T"]), int(g["T_ref"]), seed=int(g["seed_x"]), ragged=bool(g["ragged"]),
                       dur_lo=int(g["dur_lo"]), dur_hi=int(g["dur_hi"]))
    model = build_model(cfg, w, device=DEV, tensor_core_policy=policy)
    force = bool(g["force"])
    with torch.no_grad():
        wav, mel, mel_len, logd = model(dict(x), force_duration=force)  # host tensors in, like export_hifigan.py
    ref = {k: g[k] for k in ("wav", "mel", "mel_len", "log_duration")}
    if policy == 1:  # stage-wise: style vector vs the reference's, the rest vs the oracle fed the engine's style vector
        with torch.no_grad():
            style_e = model._spkemb(x["ref_mel"].to(DEV)).cpu()
            check("style_embed (TF32 speaker net)", style_e, g["style_embed"], **TC_MEL)
            owav, omel, olen, ologd, _ = zo.zerovox_forward(cfg, w, dict(x), force_duration=force, style_embed=style_e)
        ref = {"wav": owav.numpy(), "mel": omel.numpy(), "mel_len": olen.numpy(), "log_duration": ologd.numpy()}
    check("log_duration", logd, ref["log_duration"], **FP32)
    if not np.array_equal(mel_len.cpu().numpy(), ref["mel_len"]):
        # only legitimate when a predicted duration sat on a rounding boundary
        assert not force
        pytest.skip("predicted duration flipped at a rounding boundary (documented in DESIGN.md)")
    assert wav.shape == ref["wav"].shape and mel.shape == ref["mel"].shape
    check("mel", mel, ref["mel"], **tol(policy, "mel"))
    check("wav", wav, ref["wav"], **tol(policy, "wav"))
    assert float(wav.abs().max()) <= 1.0
    # batch-1 inference_ex with the stateful _min_mel_len padding (model.py:308-347)
    x1 = {k: v[:1] for k, v in x.items() if k != "phoneme_mask"}
    model._min_mel_len = int(g["ix_min_mel_len"])
    style = torch.from_numpy(g["style_embed"][:1]).to(DEV)
    with torch.no_grad():
        iwav, ilen, ilogd, imel = model.inference_ex(to_dev(x1), style_embed=style, force_duration=force)
    if ilen == int(g["ix_mel_len"]):
        assert iwav.shape[0] == ilen * cfg.hop_length and imel.shape == (cfg.n_mels, ilen)
        check("inference_ex.mel", imel, g["ix_mel"], **tol(policy, "mel"))
        check("inference_ex.wav", iwav, g["ix_wav"], **tol(policy, "wav"))
        assert model._min_mel_len == max(int(g["ix_min_mel_len"]), ilen)
        w3, l3, d3 = model.inference(to_dev(x1), style_embed=style)  # 3-tuple variant (model.py:349-351)
        assert isinstance(l3, int) and d3.shape == ilogd.shape


@pytest.mark.parametrize("policy", [0, 1])
def test_forward_matches_oracle_ragged_batch(medium, policy):
    cfg = medium.cfg
    x = zo.make_inputs(cfg, 4, 21, 64, seed=13, ragged=True)
    with torch.no_grad():
        wav, mel, mel_len, logd = medium.model(policy)(dict(x), force_duration=True)
        style_e = medium.model(policy)._spkemb(x["ref_mel"].to(DEV)).cpu() if policy else None   # see module docstring
        owav, omel, olen, ologd, st = zo.zerovox_forward(cfg, medium.w, dict(x), force_duration=True, style_embed=style_e)
    assert torch.equal(mel_len.cpu(), olen)
    check("log_duration", logd, ologd, **FP32)
    check("mel", mel, omel, **tol(policy, "mel"))
    check("wav", wav, owav, **tol(policy, "wav"))


@pytest.mark.parametrize("policy", [1])
def test_full_size_properties_config2(medium, policy):
    """BASELINE config 2 (B=32, T=128, forced durations U{2..10}): size-independent properties."""
    cfg = medium.cfg
    B, T = 32, 128
    x = zo.make_inputs(cfg, B, T, 440, seed=7)
    model = medium.model(policy)
    with torch.no_grad():
        wav, mel, mel_len, logd = model(dict(x), force_duration=True)
    L = int(mel_len.max())
    assert mel_len.cpu().tolist() == x["duration"].sum(1).tolist()       # mel_len == sum(duration)
    assert wav.shape == (B, L * cfg.hop_length) and mel.shape == (B, cfg.n_mels, L) and logd.shape == (B, T)
    assert torch.isfinite(wav).all() and torch.isfinite(mel).all()
    assert float(wav.abs().max()) <= 1.0                                  # tanh
    # utterances are independent: item 5 alone == item 5 in the batch, except where the batch padding is visible
    # (last ~14 frames: padded mel frames equal mel_linear.bias, SURVEY.md §7 quirk b)
    i = 5
    xi = {k: v[i:i + 1] for k, v in x.items()}
    with torch.no_grad():
        wav1, mel1, len1, _ = model(xi, force_duration=True)
    n = int(len1[0])
    assert n == int(mel_len[i])
    check("mel batch-invariance", mel1[0, :, :n], mel[i, :, :n], rtol=3e-3 if policy else 1e-4)   # measured 1.5e-3 * max
    keep = (n - 16) * cfg.hop_length
    check("wav batch-invariance", wav1[0, :keep], wav[i, :keep], rtol=0.0, atol=1.6e-2 if policy else 1e-4)   # measured 7.8e-3
    # speaker embedding has unit norm
    style = model._spkemb(x["ref_mel"][:4].to(DEV))
    np.testing.assert_allclose(style.norm(dim=-1).cpu().numpy(), 1.0, atol=1e-5)


def test_launches_are_counted_and_native(tiny):
    model = tiny.model(1)
    eng = model._shared_ctx.get(torch.device(DEV))
    before = eng.launch_count()
    x = zo.make_inputs(tiny.cfg, 2, 7, 16, seed=1)
    with torch.no_grad():
        model(dict(x), force_duration=True)
    assert eng.launch_count() - before > 50
    assert eng.workspace_bytes() > 0


# ---------------------------------------------------------------------------------------------- long-form (config 5)
def test_length_regulator_chunks_equal_full(medium):
    """zvx_length_regulate_chunk: the frames [f0, f0+n) of the full gather, indices bit-exact."""
    cfg = medium.cfg
    g = torch.Generator().manual_seed(4)
    B, T = 2, 300
    x = torch.randn(B, T, cfg.hidden, generator=g).to(DEV)
    dur = torch.randint(0, 9, (B, T), generator=g, dtype=torch.int32).to(DEV)
    eng = medium.model(1)._shared_ctx.get(torch.device(DEV))
    L = int(dur.sum(1).max())
    full, idx = eng.length_regulate(x, dur, L, want_index=True)
    for f0, n in ((0, 100), (100, 333), (433, L - 433), (L - 5, 40)):
        part, pidx = eng.length_regulate(x, dur, n, want_index=True, frame0=f0)
        m = min(n, L - f0)
        assert torch.equal(part[:, :m], full[:, f0:f0 + m]) and torch.equal(pidx[:, :m], idx[:, f0:f0 + m])
        if m < n:   # frames past every utterance's end: zero rows, index -1
            assert float(part[:, m:].abs().max()) == 0.0 and bool((pidx[:, m:] == -1).all())


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
def test_vocoder_chunked_equals_full(v):
    """Overlap-discard chunking with a 14-frame halo reproduces the unchunked waveform (receptive field < 14 frames)."""
    h = {"v1": zo.HifiGanConfig.v1, "v2": zo.HifiGanConfig.v2, "v3": zo.HifiGanConfig.v3}[v]()
    g = torch.Generator().manual_seed(5)
    hw = zo.make_hifigan_weights(h, g)
    mel = torch.randn((2, 80, 150), generator=g).to(DEV)
    gen = build_generator(h, {"_meldec." + k: t for k, t in hw.items()}).to(DEV)
    with torch.no_grad():
        full = gen(mel)
        for chunk in (37, 64):
            part = gen.forward_chunked(mel, chunk_frames=chunk, halo_frames=14)
            assert part.shape == full.shape
            # not bit-exact: in the polyphase kernel the order of the tensor-core tap sum depends on a sample's phase
            # inside its tile, and the tile origin moves with the chunk
            check(f"chunked({chunk}) vs full {v}", part, full, rtol=0.0, atol=5e-4)


def test_longform_decoder_attention_chunks(medium):
    """L > max_mel_len (position table recomputed, fs2.py:287-294) with the attention score budget forced small so that
    query rows are processed in several chunks: same mel as the oracle, and as the unchunked engine."""
    import zerovox_b200.engine as engine_mod
    cfg = medium.cfg
    g = torch.Generator().manual_seed(6)
    B, L = 1, 1900
    feats = torch.randn(B, L, cfg.hidden, generator=g)
    style = torch.nn.functional.normalize(torch.randn(B, 1, cfg.hidden, generator=g), dim=-1)
    mask = torch.zeros(B, L, dtype=torch.bool)
    with torch.no_grad():
        ref = zo.fs2_decoder(cfg, medium.w, feats, mask, style)
    eng = medium.model(1)._shared_ctx.get(torch.device(DEV))
    full, _ = eng.decode(feats.to(DEV), style.to(DEV), mask=mask.to(DEV), want_bcl=False)
    check("long-form mel vs oracle", full, ref, **TC_MEL)
    # the three-kernel attention (fused_attention = 0: QK^T GEMM, softmax, PV GEMM with the scores in HBM) on a second engine,
    # unchunked and with the score budget forced small
    eng2 = engine_mod.Engine(eng.cfg, torch.device(DEV))
    eng2.set_option("fused_attention", 0)
    eng2.load_weights({k: v for k, v in medium.w.items() if k.startswith("_mel_decoder.")})
    unfused, _ = eng2.decode(feats.to(DEV), style.to(DEV), mask=mask.to(DEV), want_bcl=False)
    check("three-kernel attention vs oracle", unfused, ref, **TC_MEL)
    eng2.set_option("score_workspace_bytes", 2 * 2 * 500 * 1900 * 4)   # ~500 query rows per chunk
    part, _ = eng2.decode(feats.to(DEV), style.to(DEV), mask=mask.to(DEV), want_bcl=False)
    check("chunked attention vs unchunked", part, unfused, rtol=0.0, atol=1e-5)
    # each variant of the fused kernel on its own
    for variant, name in ((2, "single-CTA"), (3, "CTA-pair")):
        eng2.set_option("fused_attention", variant)
        one, _ = eng2.decode(feats.to(DEV), style.to(DEV), mask=mask.to(DEV), want_bcl=False)
        check(f"fused attention ({name}) vs oracle", one, ref, **TC_MEL)


def test_longform_config5_properties(medium):
    """BASELINE config 5 at full size: one 4096-phoneme utterance, forced durations U{2..10} (L ~ 24.5k frames > every
    table), chunked vocoder.  Size-independent properties + the chunked vocoder against the unchunked one."""
    cfg = medium.cfg
    x = zo.make_inputs(cfg, 1, 4096, 64, seed=11)
    model = medium.model(1)
    with torch.no_grad():
        style = model._spkemb(x["ref_mel"].to(DEV))
        x1 = {k: v.to(DEV) for k, v in x.items() if k != "ref_mel"}
        wav, mel_len, logd, mel = model.inference_ex(x1, style_embed=style, force_duration=True, vocoder_chunk_frames=512)
        wav_full, mel_len2, _, _ = model.inference_ex(x1, style_embed=style, force_duration=True)
    model._min_mel_len = 689   # inference_ex grows it to the longest utterance seen (model.py:331-335); restore for other tests
    assert mel_len == mel_len2 == int(x["duration"].sum())                      # export_hifigan.py:125-128
    assert wav.shape == (mel_len * cfg.hop_length,) and mel.shape == (cfg.n_mels, mel_len)
    assert torch.isfinite(wav).all() and float(wav.abs().max()) <= 1.0
    check("config5 chunked vs full vocoder", wav, wav_full, rtol=0.0, atol=5e-4)
