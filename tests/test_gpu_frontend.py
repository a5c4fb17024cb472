"""GPU parity of the speaker-prompt front-end (csrc/frontend.cu) through the C ABI, against the
oracle (oracle/frontend_oracle.py) and the goldens the reference's own get_mel_from_wav produced.

Tolerances: the FFT runs in fp32 on the GPU and in float64 in librosa / the oracle, so a magnitude carries an absolute
error of ~1e-6 x the frame's largest bin; on the log-mel that is <= 2e-3 absolute wherever the mel value is > 1e-3 x the
frame's largest mel value (and the linear mel is compared everywhere, relative to the frame's largest);
trim bounds (integers) are exact for signals whose frames are not within 0.01 dB of the threshold."""
import os

import numpy as np
import pytest
import torch

from oracle import frontend_oracle as fo
from oracle import zerovox_oracle as zo
from zerovox_b200.frontend import MelFrontend
from zerovox_b200.synthetic import make_speech_like
from zerovox_b200.testing import build_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def fe():
    return MelFrontend(device=DEV)


def check_mel(got_BTC, ref_spec, got_energy=None, ref_energy=None):
    got = got_BTC.cpu().numpy().T                    # [80, frames] like the reference
    assert got.shape == ref_spec.shape, (got.shape, ref_spec.shape)
    lin_g, lin_r = np.exp(got.astype(np.float64)), np.exp(ref_spec.astype(np.float64))
    fmax = lin_r.max(axis=0, keepdims=True)
    e_lin = (np.abs(lin_g - lin_r) / fmax).max()
    sig = lin_r > 1e-3 * fmax
    e_log = np.abs(got - ref_spec)[sig].max()
    print(f"  frames {got.shape[1]:5d}  linear mel err / frame max {e_lin:.2e}   log-mel err (significant bins) {e_log:.2e}")
    assert e_lin < 2e-5 and e_log < 2e-3
    if got_energy is not None:
        e = np.abs(got_energy.cpu().numpy() - ref_energy).max() / ref_energy.max()
        print(f"  energy rel err {e:.2e}")
        assert e < 1e-5


@pytest.mark.parametrize("name", ["short", "prompt"])
def test_mel_matches_reference_golden(fe, golden_dir, name):
    g = np.load(os.path.join(golden_dir, "melfront.npz"))
    wav = make_speech_like(int(g[name + "_n"]), seed=int(g[name + "_seed"]))
    mel, energy = fe.mel(torch.from_numpy(wav).to(DEV), with_energy=True)
    check_mel(mel[0], g[name + "_spec"], energy[0], g[name + "_energy"])


@pytest.mark.parametrize("n", [385, 1024, 1279, 1280, 22050 * 5 + 17])
def test_mel_matches_oracle_edge_lengths(fe, n):
    # 385 = shortest legal input (reflect padding needs n > 384); 1279/1280 straddle a frame boundary
    wav = make_speech_like(n, seed=n % 97, lead=0.0, tail=0.0)
    spec, energy = fo.get_mel_from_wav(wav)
    assert fe.num_frames(n) == spec.shape[1] == n // 256
    mel, en = fe.mel(torch.from_numpy(wav).to(DEV), with_energy=True)
    check_mel(mel[0], spec, en[0], energy)


def test_mel_too_short_gives_no_frames(fe):
    assert fe.num_frames(384) == 0 and fe.num_frames(0) == 0
    mel = fe.mel(torch.zeros(300, device=DEV))
    assert mel.shape == (1, 0, 80)


def test_mel_pure_tone_and_silence(fe):
    n = 8192
    t = np.arange(n) / 22050.0
    tone = (0.5 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
    spec, _ = fo.get_mel_from_wav(tone)
    check_mel(fe.mel(torch.from_numpy(tone).to(DEV))[0], spec)
    silent = fe.mel(torch.zeros(n, device=DEV))[0].cpu().numpy()
    assert np.array_equal(silent, np.full_like(silent, np.log(np.float32(1e-5))))     # the clip floor, exactly


def test_mel_batched_windows_match_per_item(fe):
    """[B, n] rows with device-side (start, len) windows — rows are zero-filled beyond their own frame count."""
    lens = [30000, 12345, 700, 22050]
    starts = [0, 1000, 64, 3]
    n = 32000
    wavs = np.stack([make_speech_like(n, seed=10 + i, lead=0.0, tail=0.0) for i in range(len(lens))])
    F = max(l // 256 for l in lens)
    mel, en = fe.mel(torch.from_numpy(wavs).to(DEV), wav_start=torch.tensor(starts, device=DEV),
                     wav_len=torch.tensor(lens, device=DEV), n_frames=F, with_energy=True)
    for b, (s, l) in enumerate(zip(starts, lens)):
        spec, energy = fo.get_mel_from_wav(wavs[b, s:s + l])
        k = spec.shape[1]
        check_mel(mel[b, :k], spec, en[b, :k], energy)
        assert not mel[b, k:].any() and not en[b, k:].any()


def test_trim_matches_oracle(fe):
    n = 44223
    wavs = np.stack([make_speech_like(n, seed=s, lead=l, tail=t) for s, l, t in ((2, .12, .10), (3, .0, .3), (4, .25, .0))]
                    + [np.full(n, 0.3, dtype=np.float32), np.zeros(n, dtype=np.float32)])
    lens = [n, n, 30000, n, 5000]
    start, length, hs, hl = fe.trim(torch.from_numpy(wavs).to(DEV), wav_len=torch.tensor(lens, device=DEV))
    for b in range(len(lens)):
        _, (a, e) = fo.trim(wavs[b, :lens[b]])
        assert (hs[b], hs[b] + hl[b]) == (a, e), (b, hs[b], hl[b], a, e)
    assert start.tolist() == hs and length.tolist() == hl
    # other trim parameters
    _, _, hs, hl = fe.trim(torch.from_numpy(wavs[:1]).to(DEV), top_db=20.0, frame_length=1024, hop_length=256)
    _, (a, e) = fo.trim(wavs[0], top_db=20.0, frame_length=1024, hop_length=256)
    assert (hs[0], hs[0] + hl[0]) == (a, e)


@pytest.mark.parametrize("sr_in,sr_out", [(24000, 22050), (16000, 22050), (44100, 22050), (22050, 24000)])
def test_resample_matches_oracle(fe, sr_in, sr_out):
    """zvx_resample (the `sr=` of librosa.load, synthesize.py:113-121) against the float64 oracle of the same filter: fp32
    accumulation of ~140 taps -> 3e-6; ragged rows are zero-extended; and against the analytic band-limited answer."""
    n = 30011
    rng = np.random.default_rng(1)
    wavs = np.stack([make_speech_like(n, seed=3), rng.normal(0, 0.2, n).astype(np.float32),
                     np.sin(2 * np.pi * 1234.5 * np.arange(n) / sr_in).astype(np.float32)])
    lens = [n, 20000, n]
    out = fe.resample(torch.from_numpy(wavs).to(DEV), sr_in, sr_out, wav_len=torch.tensor(lens, device=DEV)).cpu().numpy()
    assert out.shape == (3, -(-n * sr_out // sr_in))
    for b in range(3):
        ref = fo.resample(wavs[b, :lens[b]], sr_in, sr_out)
        assert np.abs(out[b, :len(ref)] - ref).max() < 3e-6
        assert np.abs(out[b, len(ref) + 80:]).max() < 1e-6 if len(ref) + 80 < out.shape[1] else True
    t_out = np.arange(out.shape[1]) / sr_out
    assert np.abs(out[2] - np.sin(2 * np.pi * 1234.5 * t_out))[200:-200].max() < 2e-5
    same = fe.resample(torch.from_numpy(wavs).to(DEV), sr_out, sr_out)
    assert same.shape == wavs.shape and torch.equal(same.cpu(), torch.from_numpy(wavs))


def test_speaker_prompt_mel_and_speaker_embed_mirror():
    """ZeroVoxTTS.speaker_embed (synthesize.py:123-143) through the mirror: trim -> mel -> `_spkemb`, against the
    oracle composition (speaker net in fp32 FMA, policy 0, so the only differences are the front-end's)."""
    from zerovox_b200.tts.symbols import Symbols
    from zerovox_b200.tts.synthesize import ZeroVoxTTS
    cfg = zo.ZeroVoxConfig.tiny()
    w = zo.make_weights(cfg, seed=0)
    model = build_model(cfg, w, device=DEV, tensor_core_policy=0)
    tts = ZeroVoxTTS(language="en", syms=Symbols(cfg.phones, cfg.puncts), checkpoint=None, meldec_model=None,
                     hop_length=256, sampling_rate=22050, n_mel_channels=80, fft_size=1024, win_length=1024, mel_fmin=0,
                     mel_fmax=8000, infer_device=DEV, model=model)
    wav = make_speech_like(22050 * 2, seed=5)
    ref_mel = fo.speaker_prompt_mel(wav)
    got_mel = tts._frontend.speaker_prompt_mel(torch.from_numpy(wav).to(DEV))
    check_mel(got_mel[0], ref_mel[0].T)
    style = tts.speaker_embed(wav)                                  # numpy in, like the reference's caller
    with torch.no_grad():
        ref = zo.speaker_embed(cfg, w, torch.from_numpy(ref_mel))
    assert style.shape == ref.shape
    err = (style.cpu() - ref).abs().max().item()
    print(f"  style |diff| {err:.2e}  (unit-norm vector of {ref.shape[-1]})")
    assert err < 2e-4 and abs(style.norm().item() - 1.0) < 1e-5
    assert tts.transcript2phonemids("this is a test.") == fo.transcript2phonemids(fo.Symbols(cfg.phones, cfg.puncts),
                                                                                  "this is a test.")
    # a 24 kHz prompt (every packaged prompt of the reference): resampled on the GPU first, like librosa.load(sr=22050)
    wav24 = make_speech_like(24000 * 2, seed=6, sr=24000)
    style24 = tts.speaker_embed(wav24, sampling_rate=24000)
    with torch.no_grad():
        ref24 = zo.speaker_embed(cfg, w, torch.from_numpy(fo.speaker_prompt_mel(fo.resample(wav24, 24000, 22050))))
    assert (style24.cpu() - ref24).abs().max().item() < 2e-4


def test_get_mel_from_wav_mirror_signature():
    from zerovox_b200.tts.mels import get_mel_from_wav
    wav = make_speech_like(6000, seed=8)
    spec, energy = get_mel_from_wav(audio=wav, sampling_rate=22050, fft_size=1024, hop_size=256, win_length=1024,
                                    num_mels=80, fmin=0, fmax=8000)
    ospec, oenergy = fo.get_mel_from_wav(wav)
    assert isinstance(spec, np.ndarray) and spec.shape == ospec.shape and energy.shape == oenergy.shape
    check_mel(torch.from_numpy(spec.T.copy()), ospec)


def test_frontend_rejects_unsupported_config_and_cpu_tensors(fe):
    with pytest.raises(RuntimeError, match="fft_size"):
        MelFrontend(fft_size=2048, win_length=2048, device=DEV)
    with pytest.raises(RuntimeError, match="fmax"):
        MelFrontend(fmax=20000, device=DEV)
    with pytest.raises(RuntimeError, match="no CPU path"):
        fe.mel(torch.zeros(4000))


def test_tts_ex_mirror_end_to_end_text_to_waveform():
    """ZeroVoxTTS.tts_ex (synthesize.py:215-239) through the mirror: normalised text -> native tokeniser -> inference_ex with
    forced durations -> waveform, against the oracle fed the same ids (policy 0: fp32 FMA throughout)."""
    from zerovox_b200.tts.symbols import Symbols
    from zerovox_b200.tts.synthesize import ZeroVoxTTS

    class Lower:                                     # stands in for ZeroVoxNormalizer.normalize (normalize.py, out of scope)
        def normalize(self, text):
            return text.lower(), None

    cfg = zo.ZeroVoxConfig.tiny()
    w = zo.make_weights(cfg, seed=0)
    model = build_model(cfg, w, device=DEV, tensor_core_policy=0)
    tts = ZeroVoxTTS(language="en", syms=Symbols(cfg.phones, cfg.puncts), checkpoint=None, meldec_model=None,
                     hop_length=256, sampling_rate=22050, n_mel_channels=80, fft_size=1024, win_length=1024, mel_fmin=0,
                     mel_fmax=8000, infer_device=DEV, model=model, normalizer=Lower())
    text = "  Hello, World! This is it.  "
    ph, pu = fo.transcript2phonemids(fo.Symbols(cfg.phones, cfg.puncts), text.strip().lower())
    assert tts.text2phonemeids(text.strip()) == (ph, pu) and len(ph) == 18
    dur = [(3 * i) % 5 + 1 for i in range(len(ph))]
    style = tts.speaker_embed(make_speech_like(22050, seed=6))
    wav, phoneme, length, mel = tts.tts_ex(text, style, duration=dur)
    x = {"phoneme": torch.tensor([ph], dtype=torch.int32), "puncts": torch.tensor([pu], dtype=torch.int32),
         "duration": torch.tensor([dur], dtype=torch.int32)}
    with torch.no_grad():
        owav, olen, _, omel, _ = zo.zerovox_inference_ex(cfg, w, x, style.cpu(), force_duration=True)
    assert length == olen == sum(dur) and phoneme.cpu().tolist() == [ph]
    assert wav.shape == (olen * 256,) and mel.shape == (80, olen)
    e_mel = np.abs(mel - omel.numpy()).max() / np.abs(omel.numpy()).max()
    e_wav = np.abs(wav - owav.numpy()).max()
    print(f"  tts_ex: mel rel err {e_mel:.2e}, wav abs err {e_wav:.2e}")
    assert e_mel < 5e-4 and e_wav < 5e-4
    # empty text: the reference's early return (synthesize.py:221-222)
    w0, p0, l0, m0 = tts.tts_ex("   ", style)
    assert l0 == 0 and w0.shape == (1, 1) and p0.shape == (1, 1)


def test_load_model_directory_then_speak(tmp_path, monkeypatch):
    """ZeroVoxTTS.load_model on the reference's on-disk layout (synthesize.py:275-328): a model directory with modelcfg.yaml +
    checkpoints/*.ckpt (Lightning format) and a vocoder repo in the cache layout with weight-normed generator.ckpt; then the
    demo flow (zerovox/demo.py:103-115): speaker_embed(prompt) -> tts(text) -> waveform, against the oracle."""
    import json
    import yaml
    from zerovox_b200.testing import zerovox_kwargs
    from zerovox_b200.tts import AttrDict, Generator
    from zerovox_b200.tts.synthesize import ZeroVoxTTS

    class Lower:
        def normalize(self, text):
            return text.lower(), None

    cfg = zo.ZeroVoxConfig.tiny()
    w = zo.make_weights(cfg, seed=3)
    name = "zerovox-hifigan-test"
    repo = tmp_path / "model_repo" / name
    repo.mkdir(parents=True)
    (repo / "config.json").write_text(json.dumps(cfg.hifigan.as_json_dict()))
    raw_gen = Generator(AttrDict(cfg.hifigan.as_json_dict()))          # weight-normed, like upstream generator.ckpt
    torch.save({"generator": raw_gen.state_dict()}, repo / "generator.ckpt")
    monkeypatch.setenv("CACHED_PATH_ZEROVOX", str(tmp_path))
    mdir = tmp_path / "tts_en_test"
    (mdir / "checkpoints").mkdir(parents=True)
    (mdir / "modelcfg.yaml").write_text(yaml.safe_dump({
        "lang": ["en"], "model": {"phones": cfg.phones, "puncts": cfg.puncts},
        "audio": {"sampling_rate": 22050, "fft_size": 1024, "fmax": 8000, "fmin": 0, "win_length": 1024, "num_mels": 80,
                  "hop_size": 256}}))
    hp = {k: v for k, v in zerovox_kwargs(cfg).items() if k != "meldec_model"}
    sd = {k: v for k, v in w.items() if not k.startswith("_meldec.")}
    torch.save({"hyper_parameters": hp, "state_dict": sd}, mdir / "checkpoints" / "epoch=0003.ckpt")

    modelcfg, tts = ZeroVoxTTS.load_model(str(mdir), meldec_model=name, infer_device=DEV, normalizer=Lower())
    assert modelcfg["audio"]["hop_size"] == 256
    tts._model._shared_ctx.tensor_core_policy = 0                      # fp32 FMA: tight comparison
    tts._model._shared_ctx.mark_stale()
    prompt = make_speech_like(22050, seed=9)
    style = tts.speaker_embed(prompt)
    wav_p, phoneme_p, length_p = tts.tts("Hello there, world.", style)          # the demo call (predicted durations)
    assert wav_p.shape == (length_p * 256,) and phoneme_p.shape[1] == 15
    # oracle with the same weights: vocoder = the folded (remove_weight_norm) form of the raw generator
    wo = dict(sd)
    wo.update({"_meldec." + k: v for k, v in raw_gen._engine_state_dict().items()})
    ph, pu = fo.transcript2phonemids(fo.Symbols(cfg.phones, cfg.puncts), "hello there, world.")
    assert len(ph) == 15
    dur = [(5 * i) % 4 + 2 for i in range(len(ph))]                               # forced: no rounding boundary in the compare
    wav, phoneme, length, mel = tts.tts_ex("Hello there, world.", style, duration=dur)
    x = {"phoneme": torch.tensor([ph], dtype=torch.int32), "puncts": torch.tensor([pu], dtype=torch.int32),
         "duration": torch.tensor([dur], dtype=torch.int32)}
    with torch.no_grad():
        ostyle = zo.speaker_embed(cfg, wo, torch.from_numpy(fo.speaker_prompt_mel(prompt)))
        owav, olen, _, omel, _ = zo.zerovox_inference_ex(cfg, wo, x, style.cpu(), force_duration=True)
    assert (style.cpu() - ostyle).abs().max().item() < 2e-4
    assert length == olen == sum(dur) and phoneme.cpu().tolist() == [ph] and wav.shape == (olen * 256,)
    e_wav = np.abs(wav - owav.numpy()).max()
    e_mel = np.abs(mel - omel.numpy()).max() / np.abs(omel.numpy()).max()
    print(f"  load_model -> tts_ex: {olen} frames, mel rel err {e_mel:.2e}, wav err {e_wav:.2e}")
    assert e_wav < 5e-4 and e_mel < 5e-4
    with pytest.raises(FileNotFoundError):
        ZeroVoxTTS.load_model("no-such-model", meldec_model=name, infer_device=DEV)
