"""CPU: the front-end oracle (oracle/frontend_oracle.py) against the committed goldens (produced by the reference's own
get_mel_from_wav / transcript2phonemids — see oracle/make_goldens_frontend.py) and against independent implementations
(scipy window, torch.stft, torchaudio Slaney filterbank); the native tokeniser / collator (host C code in the library,
no GPU involved) against the same goldens and the oracle."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import frontend_oracle as fo
from zerovox_b200.frontend import Tokeniser
from zerovox_b200.synthetic import make_speech_like

PHONES_EN = "'-abcdefghijklmnopqrstuvwxyz"
PUNCTS_EN = " ,.;:-!?\""


# ------------------------------------------------------------------------------------------------ mel front-end oracle
def test_hann_window_is_scipys():
    from scipy.signal import get_window
    for n in (1024, 800, 17):
        assert np.allclose(fo.hann_periodic(n), get_window("hann", n, fftbins=True), rtol=0, atol=1e-15)


def test_mel_filterbank_matches_torchaudio_slaney():
    import torchaudio
    for sr, n_fft, n_mels, fmin, fmax in ((22050, 1024, 80, 0, 8000), (24000, 1024, 100, 50, 12000)):
        ref = torchaudio.functional.melscale_fbanks(n_freqs=1 + n_fft // 2, f_min=float(fmin), f_max=float(fmax),
                                                    n_mels=n_mels, sample_rate=sr, norm="slaney", mel_scale="slaney").T
        got = fo.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
        assert got.dtype == np.float32 and got.shape == (n_mels, 1 + n_fft // 2)
        # torchaudio builds its table in float32 throughout, librosa (and the restatement) in float64 then rounds
        assert np.abs(got - ref.numpy()).max() < 5e-6 * np.abs(got).max()
        assert (got >= 0).all() and (got.sum(axis=1) > 0).all()       # no empty filters at these settings


def test_stft_magnitude_matches_torch_stft():
    y = make_speech_like(9000, seed=3)
    mag = fo.stft_magnitude(y, 1024, 256, 1024)
    ref = torch.stft(torch.from_numpy(y).double(), 1024, hop_length=256, win_length=1024,
                     window=torch.hann_window(1024, periodic=True, dtype=torch.float64), center=False,
                     return_complex=True).abs().numpy()
    assert mag.shape == ref.shape == (513, 1 + (9000 - 1024) // 256)
    assert np.abs(mag - ref).max() < 1e-6 * ref.max()


@pytest.mark.parametrize("name", ["short", "prompt"])
def test_get_mel_from_wav_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "melfront.npz"))
    wav = make_speech_like(int(g[name + "_n"]), seed=int(g[name + "_seed"]))
    spec, energy = fo.get_mel_from_wav(wav)
    assert spec.shape == g[name + "_spec"].shape == (80, len(wav) // 256)
    assert np.abs(spec - g[name + "_spec"]).max() < 2e-4            # log-mel, absolute
    assert np.abs(energy - g[name + "_energy"]).max() < 1e-5 * g[name + "_energy"].max()


def test_trim_invariants():
    # librosa.effects.trim is restated, not pinned (librosa absent): check what its definition guarantees
    n = 44223
    wav = make_speech_like(n, seed=2)            # ~12 % near-silent lead-in, ~10 % tail at -80 dB
    y, (a, b) = fo.trim(wav)
    assert a % 512 == 0 and (b % 512 == 0 or b == n) and 0 < a < b < n
    assert abs(a - int(0.12 * n)) <= 1024 and abs(b - (n - int(0.10 * n))) <= 1536    # within ~a frame of the gates
    assert np.array_equal(y, wav[a:b])
    loud = np.ones(5000, dtype=np.float32) * 0.3
    assert fo.trim(loud)[1] == (0, 5000)                                # nothing to cut
    assert fo.trim(np.zeros(3000, dtype=np.float32))[1] == (0, 3000)    # all frames equal the (clamped) reference level
    assert fo.trim(wav, top_db=200.0)[1] == (0, n)


def test_trim_two_independent_formulations_agree():
    """librosa is absent, so `trim` cannot be pinned to librosa itself; it is cross-checked against a second statement of the
    same published rule written differently (float64 cumulative-sum frame power, power-ratio threshold instead of a dB
    difference) on 40 signals: speech-like envelopes with varying lead / tail, noise floors around the threshold, clicks."""
    rng = np.random.default_rng(0)
    checked = 0
    for s_ in range(40):
        n = int(rng.integers(6000, 60000))
        w = make_speech_like(n, seed=s_, lead=float(rng.uniform(0, .3)), tail=float(rng.uniform(0, .3)))
        if s_ % 4 == 1:
            w = w + rng.normal(0, 10 ** rng.uniform(-4.5, -2.0), n).astype(np.float32)    # noise floor near / above -40 dB
        if s_ % 4 == 2:
            w[int(rng.integers(0, n))] += 0.9                                               # a click
        for top_db, fl, hop in ((40.0, 2048, 512), (20.0, 1024, 256)):
            a = fo.trim(w, top_db=top_db, frame_length=fl, hop_length=hop)[1]
            b = fo.trim_bounds_independent(w, top_db=top_db, frame_length=fl, hop_length=hop)
            if a != b:   # only legitimate when a frame's power sits within float32 rounding of the threshold
                yp = np.pad(w.astype(np.float64), (fl // 2, fl // 2))
                pw = np.array([np.mean(yp[i * hop:i * hop + fl] ** 2) for i in range(1 + (len(yp) - fl) // hop)])
                thr = max(pw.max(), 1e-10) * 10 ** (-top_db / 10)
                assert np.min(np.abs(pw / thr - 1.0)) < 1e-5, (s_, a, b)
            checked += 1
    assert checked == 80


def test_resample_oracle_is_band_limited_interpolation():
    """The `sr=` conversion of librosa.load (synthesize.py:113-121).  soxr is absent, so parity with librosa's soxr_hq is a
    stated tolerance: both are band-limited interpolators, and for a band-limited input the exact answer is known — sinusoids
    below 0.85 x the output Nyquist frequency, sampled analytically at the output rate."""
    from scipy.signal import resample_poly
    for sr_in, sr_out in ((24000, 22050), (16000, 22050), (44100, 22050)):
        n = sr_in
        t_in = np.arange(n) / sr_in
        freqs = [110.0, 997.0, 3501.0, 0.70 * min(sr_in, sr_out) / 2, 0.85 * min(sr_in, sr_out) / 2]
        x = sum(np.sin(2 * np.pi * f * t_in + i) / len(freqs) for i, f in enumerate(freqs))
        y = fo.resample(x.astype(np.float32), sr_in, sr_out)
        assert len(y) == -(-n * sr_out // sr_in)                       # ceil(n * ratio), librosa.resample
        t_out = np.arange(len(y)) / sr_out
        ref = sum(np.sin(2 * np.pi * f * t_out + i) / len(freqs) for i, f in enumerate(freqs))
        mid = slice(200, -200)                                          # the truncated ends are not band-limited
        assert np.abs(y[mid] - ref[mid]).max() < 2e-6, (sr_in, sr_out, np.abs(y[mid] - ref[mid]).max())
        g = np.gcd(sr_in, sr_out)
        z = resample_poly(x, sr_out // g, sr_in // g)                   # another band-limited resampler (Kaiser beta = 5 FIR)
        assert np.abs(z[mid] - y[mid]).max() < 5e-3
    assert len(fo.resample(np.zeros(0, dtype=np.float32), 24000, 22050)) == 0


# ------------------------------------------------------------------------------------------------ tokeniser
def _cases(golden_dir):
    with open(os.path.join(golden_dir, "tokeniser.json"), encoding="utf-8") as f:
        return json.load(f)


def test_tokeniser_oracle_matches_reference_golden(golden_dir):
    for c in _cases(golden_dir):
        ph, pu = fo.transcript2phonemids(fo.Symbols(c["phones"], c["puncts"]), c["text"])
        assert (ph, pu) == (c["phone_ids"], c["punct_ids"]), c["text"]


def test_native_tokeniser_matches_reference_golden(golden_dir):
    toks = {}
    for c in _cases(golden_dir):
        key = (c["phones"], c["puncts"])
        tok = toks.setdefault(key, Tokeniser(*key))
        assert tok.transcript2phonemids(c["text"]) == (c["phone_ids"], c["punct_ids"]), c["text"]


def test_native_symbols_counts():
    tok = Tokeniser(PHONES_EN, PUNCTS_EN)
    sym = fo.Symbols(PHONES_EN, PUNCTS_EN)
    assert tok.num_phones == sym.num_phones == 28 and tok.num_puncts == sym.num_puncts == 10
    dup = Tokeniser("abca", " ,, ")        # duplicates: dict semantics, the later index wins (symbols.py:11-22)
    sdup = fo.Symbols("abca", " ,, ")
    assert dup.num_phones == sdup.num_phones == 3 and dup.num_puncts == sdup.num_puncts == 3
    assert dup.transcript2phonemids("a, b") == fo.transcript2phonemids(sdup, "a, b") == ([3, 1], [4, 0])


def test_native_tokeniser_random_strings_match_oracle():
    rng = random.Random(5)
    alphabet = PHONES_EN + PUNCTS_EN + "ABZ019#äß€𝄞\t\n" + "   ,,.."
    tok, sym = Tokeniser(PHONES_EN, PUNCTS_EN), fo.Symbols(PHONES_EN, PUNCTS_EN)
    for _ in range(400):
        s = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 60)))
        assert tok.transcript2phonemids(s) == fo.transcript2phonemids(sym, s), repr(s)
    long = "".join(rng.choice(alphabet) for _ in range(20000))
    assert tok.transcript2phonemids(long) == fo.transcript2phonemids(sym, long)


def test_blank_without_blank_punct_raises_like_reference():
    tok, sym = Tokeniser("ab", ",."), fo.Symbols("ab", ",.")
    with pytest.raises(KeyError):
        fo.transcript2phonemids(sym, "a b")
    with pytest.raises(KeyError):
        tok.transcript2phonemids("a b")
    assert tok.transcript2phonemids("a,b") == ([0, 1], [1, 0])


def test_collate_matches_oracle_and_pad_sequence():
    from torch.nn.utils.rnn import pad_sequence
    rng = random.Random(9)
    tok = Tokeniser(PHONES_EN, PUNCTS_EN)
    for B in (1, 3, 17):
        seqs = [[rng.randint(0, 27) for _ in range(rng.randint(0 if B > 1 else 1, 40))] for _ in range(B)]
        pus = [[rng.randint(0, 9) for _ in s] for s in seqs]
        ph, pu, mask, lens = tok.collate(seqs, pus)
        oph, opu, omask, olens = fo.collate(seqs, pus)
        assert np.array_equal(ph.numpy(), oph) and np.array_equal(pu.numpy(), opu)
        assert np.array_equal(mask.numpy(), omask) and np.array_equal(lens.numpy(), olens)
        ref = pad_sequence([torch.tensor(s, dtype=torch.int32) for s in seqs], batch_first=True)   # data.py:59
        if ref.numel():
            assert torch.equal(ph, ref)
    ph, pu, mask, lens = tok.collate([[], []], [[], []])        # all-empty batch
    assert ph.shape == (2, 0) and mask.shape == (2, 0)
    with pytest.raises(ValueError):
        tok.collate([[1, 2]], [[1]])


def test_native_tokeniser_property_any_unicode():
    """Property test (hypothesis): for ANY string and ANY small vocabulary the native tokeniser equals the restatement of
    the reference loop — code points, not bytes; invalid characters skipped; output lengths equal."""
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies
    chars = st.characters(blacklist_categories=("Cs",), blacklist_characters="\x00")

    @hyp.settings(max_examples=300, deadline=None)
    @hyp.given(phones=st.text(chars, min_size=1, max_size=12), puncts=st.text(chars, min_size=0, max_size=6),
               text=st.text(chars, max_size=80))
    def prop(phones, puncts, text):
        puncts = " " + puncts                      # every shipped config lists the blank; without it the reference raises
        tok, sym = Tokeniser(phones, puncts), fo.Symbols(phones, puncts)
        body = text + phones[:3] + puncts[:2] + text[::-1]
        ph, pu = tok.transcript2phonemids(body)
        assert (ph, pu) == fo.transcript2phonemids(sym, body)
        assert len(ph) == len(pu) and tok.num_phones == sym.num_phones and tok.num_puncts == sym.num_puncts

    prop()
