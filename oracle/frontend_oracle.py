"""CPU oracle for the callers either side of the hot path: the speaker-prompt
front-end (silence trim + log-mel spectrogram) and the tokeniser / padding collator.

TEST INFRASTRUCTURE ONLY — the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import it.  Nothing under ``zerovox_b200/`` does.

Reference call sites (gooofy/zerovox @ 56a4316):
  * ``ZeroVoxTTS.speaker_embed``      zerovox/tts/synthesize.py:123-143  (trim -> get_mel_from_wav -> _spkemb)
  * ``get_mel_from_wav``              zerovox/tts/mels.py:356-394
  * ``ZeroVoxTTS.transcript2phonemids`` zerovox/tts/synthesize.py:145-190
  * ``Symbols``                       zerovox/tts/symbols.py:2-49
  * ``LJSpeechDataModule.collate_fn`` zerovox/tts/data.py:54-78 (pad_sequence + get_mask_from_lengths, fs2.py:565-573)

Pinning status
  * tokeniser / Symbols / collator: pinned to the reference's own code — ``oracle/make_goldens_frontend.py`` imports
    ``zerovox.tts.symbols`` and ``zerovox.tts.synthesize`` (absent GUI / audio packages stubbed) from /root/reference
    and writes ``tests/golden/tokeniser.npz``.
  * mel spectrogram: the arithmetic lives in the third-party dependency **librosa (pyproject.toml:34,
    ``librosa>=0.10.2``)**, which is absent from this image and from /root/reference.  Its published algorithm
    (``librosa.stft`` with a periodic Hann window, ``center=False``; ``librosa.filters.mel`` Slaney scale + Slaney
    area normalisation; ``np.abs``; ``np.dot``) is restated here and pinned against independent implementations that
    ARE present: ``scipy.signal.get_window`` (the very function librosa calls for its window), ``torch.stft`` (the
    formulation the reference keeps commented out at mels.py:330-343, same parameters) and
    ``torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='slaney')`` (documented as librosa-equivalent) —
    see tests/test_frontend_oracle.py.
  * ``librosa.effects.trim``: restated from librosa 0.10.2 (effects.py ``trim`` / ``_signal_to_frame_nonsilent``,
    feature/spectral.py ``rms``, core/spectrum.py ``amplitude_to_db`` / ``power_to_db``); **parity unpinned** — no
    implementation of it is available offline.  Only its invariants are tested.
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------------------- mel front-end
def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True) — what librosa.stft(window='hann') builds (float64)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def _hz_to_mel(f: float) -> float:
    """librosa.hz_to_mel(htk=False): linear below 1 kHz (200/3 Hz per mel), logarithmic above."""
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f >= min_log_hz:
        return min_log_mel + np.log(f / min_log_hz) / logstep
    return f / f_sp


def _mel_to_hz(m: np.ndarray) -> np.ndarray:
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    f = f_sp * m
    log_t = m >= min_log_mel
    f[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
    return f


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, norm='slaney', float32) —
    the call of mels.py:378.  Returns [n_mels, 1 + n_fft//2] float32."""
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(float(fmin)), _hz_to_mel(float(fmax)), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))       # stored as float32
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]                                # float64 product, rounded to float32
    return weights


def stft_magnitude(y: np.ndarray, n_fft: int, hop: int, win_length: int) -> np.ndarray:
    """np.abs(librosa.stft(y, n_fft, hop_length, win_length, window='hann', center=False)) — mels.py:387-389.
    librosa multiplies the float64 window into the float32 frames (so the FFT runs in double) and stores
    complex64; np.abs of that is float32.  Returns [1 + n_fft//2, n_frames] float32."""
    win = hann_periodic(win_length)
    if win_length < n_fft:                                          # librosa.util.pad_center
        lpad = (n_fft - win_length) // 2
        win = np.pad(win, (lpad, n_fft - win_length - lpad))
    n_frames = 1 + (len(y) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    frames = y[idx]                                                 # [n_fft, n_frames] float32
    spec = np.fft.rfft(win[:, None] * frames, axis=0).astype(np.complex64)
    return np.abs(spec)


def get_mel_from_wav(audio: np.ndarray, sampling_rate=22050, fft_size=1024, hop_size=256, win_length=1024,
                     num_mels=80, fmin=0, fmax=8000):
    """mels.py:356-394: reflect-pad (fft-hop)//2, STFT magnitude, mel_basis @ mag, log(clip(., 1e-5)), and the
    per-frame spectral energy ||mag||_2.  Returns (spec [num_mels, n_frames] f32, energy [n_frames] f32)."""
    audio = np.asarray(audio, dtype=np.float32)
    basis = mel_filterbank(sampling_rate, fft_size, num_mels, fmin, fmax)
    padding = (fft_size - hop_size) // 2
    audio_padded = np.pad(audio, (padding, padding), mode="reflect")
    mag = stft_magnitude(audio_padded, fft_size, hop_size, win_length)
    spec = np.dot(basis, mag)
    spec = np.log(np.clip(spec, a_min=1e-5, a_max=None))            # dynamic_range_compression_numpy, mels.py:350-351
    energy = np.linalg.norm(mag, axis=0)
    return spec.astype(np.float32), energy.astype(np.float32)


def trim(y: np.ndarray, top_db: float = 40.0, frame_length: int = 2048, hop_length: int = 512):
    """librosa.effects.trim(y, top_db=40) — synthesize.py:126 (librosa 0.10.2 semantics; PARITY UNPINNED, see the
    module docstring).  Frame RMS over zero-padded centred frames, dB relative to the loudest frame, first/last
    frame above -top_db.  Returns (y[start:end], (start, end))."""
    y = np.asarray(y, dtype=np.float32)
    yp = np.pad(y, (frame_length // 2, frame_length // 2), mode="constant")
    n_frames = 1 + (len(yp) - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n_frames)[None, :]
    power = np.mean(np.square(yp[idx]), axis=0, dtype=np.float32)   # librosa.feature.rms (float32)
    rms = np.sqrt(power)
    ref = np.max(rms)
    amin = np.float32(1e-10)                                         # amplitude_to_db: amin = 1e-5 squared
    db = (np.float32(10.0) * np.log10(np.maximum(amin, np.square(rms)))
          - np.float32(10.0) * np.log10(np.maximum(amin, np.square(ref))))
    nonzero = np.flatnonzero(db > -top_db)
    if nonzero.size > 0:
        start = int(nonzero[0]) * hop_length
        end = min(len(y), (int(nonzero[-1]) + 1) * hop_length)
    else:
        start, end = 0, 0
    return y[start:end], (start, end)


def trim_bounds_independent(y: np.ndarray, top_db: float = 40.0, frame_length: int = 2048, hop_length: int = 512):
    """A second, independently formulated statement of librosa.effects.trim's published rule, used to cross-check `trim`
    (oracle/make_goldens_frontend.py): frame POWER by a cumulative sum in float64 instead of gathering frames, and the
    threshold as a power ratio  P_frame > P_max * 10^(-top_db/10)  instead of a difference of decibels.  Frames whose power sits
    within a float32 rounding of the threshold may legitimately differ between the two; the generator reports them."""
    y = np.asarray(y, dtype=np.float64)
    yp = np.pad(y, (frame_length // 2, frame_length // 2))
    cs = np.concatenate([[0.0], np.cumsum(yp * yp)])
    n_frames = 1 + (len(yp) - frame_length) // hop_length
    starts = hop_length * np.arange(n_frames)
    power = (cs[starts + frame_length] - cs[starts]) / frame_length
    pmax = max(power.max(), 1e-10)
    loud = np.flatnonzero(np.maximum(power, 1e-10) > pmax * 10.0 ** (-top_db / 10.0))
    if loud.size == 0:
        return 0, 0
    return int(loud[0]) * hop_length, min(len(y), (int(loud[-1]) + 1) * hop_length)


def resample(y: np.ndarray, sr_in: int, sr_out: int) -> np.ndarray:
    """The `sr=` conversion of librosa.load (synthesize.py:113-121) as band-limited interpolation with the Kaiser-windowed sinc
    of resampy's published "kaiser_best" design (64 zero crossings, beta 14.769656459379492, roll-off 0.9475937167399596),
    evaluated directly in float64 — the arithmetic csrc/frontend.cu tabulates per rational phase.  Output length
    ceil(n * sr_out / sr_in) (librosa.resample).  PARITY with librosa's default soxr_hq: unpinned (soxr is absent); both are
    > 100 dB band-limited interpolators, tests bound the difference against analytic band-limited signals instead."""
    import math
    y = np.asarray(y, dtype=np.float64)
    g = math.gcd(int(sr_in), int(sr_out))
    up, down = sr_out // g, sr_in // g
    num_zeros, beta, rolloff = 64.0, 14.769656459379492, 0.9475937167399596
    scale = min(1.0, sr_out / sr_in)
    kh = int(math.ceil(num_zeros / scale))
    n_out = -(-len(y) * sr_out // sr_in)
    m = np.arange(n_out, dtype=np.int64)
    i0 = (m * down) // up
    frac = ((m * down) % up) / up
    out = np.zeros(n_out)
    i0b = np.i0(beta)
    for j in range(2 * kh):
        k = i0 - kh + 1 + j
        u = (frac + kh - 1 - j) * scale
        w = np.where(np.abs(u) < num_zeros,
                     scale * rolloff * np.sinc(rolloff * u) * np.i0(beta * np.sqrt(np.clip(1.0 - (u / num_zeros) ** 2, 0.0, None))) / i0b, 0.0)
        ok = (k >= 0) & (k < len(y))
        out[ok] += w[ok] * y[k[ok]]
    return out.astype(np.float32)


def speaker_prompt_mel(wav: np.ndarray, **mel_kwargs) -> np.ndarray:
    """synthesize.py:123-138: trim -> get_mel_from_wav -> [1, n_frames, num_mels] (the `_spkemb` input)."""
    wav, _ = trim(wav, top_db=40)
    spec, _ = get_mel_from_wav(wav, **mel_kwargs)
    return np.array([spec.T], dtype=np.float32)


# ---------------------------------------------------------------------------------------------------- tokeniser
NO_PUNCT = "_NP_"   # symbols.py:4


class Symbols:
    """symbols.py:2-49 — phone / punct vocabularies; punct id 0 is the reserved `_NP_`, later duplicates win."""

    def __init__(self, phones: str, puncts: str):
        self.phonemap = {p: i for i, p in enumerate(phones)}
        self.punctmap = {NO_PUNCT: 0}
        for i, p in enumerate(puncts):
            self.punctmap[p] = i + 1

    @property
    def num_phones(self):
        return len(self.phonemap)

    @property
    def num_puncts(self):
        return len(self.punctmap)


def transcript2phonemids(sym: Symbols, transcript: str):
    """synthesize.py:145-190.  Runs of blanks / punctuation collapse to the largest punct id seen in the run and
    overwrite the punct slot of the PRECEDING phone; characters that are neither are skipped; a phone resets the
    running punct to 0.  (The run's punct is never reset at the start of a run: it carries over from an earlier
    run only until a phone intervenes — which always happens, so it is per run.)"""
    phones, puncts = [], []
    punct = 0
    i, n = 0, len(transcript)
    while i < n:
        p = transcript[i]
        if p == " " or p in sym.punctmap:
            while i < n and (transcript[i] == " " or transcript[i] in sym.punctmap):
                # synthesize.py:158, 168: encode_punct(' ') raises KeyError when ' ' is not a configured punct;
                # every shipped config lists it, the restatement keeps the lookup strict.
                punct = max(punct, sym.punctmap[transcript[i]])
                i += 1
            if puncts:
                puncts[-1] = punct
            continue
        if p not in sym.phonemap:
            i += 1
            continue
        punct = 0
        phones.append(sym.phonemap[p])
        puncts.append(punct)
        i += 1
    return phones, puncts


def collate(phone_seqs, punct_seqs):
    """data.py:56-60, 82-83 + fs2.py:565-573: zero-pad to the batch maximum (pad_sequence, batch_first) and build
    `phoneme_mask` = position >= length.  Returns (phoneme i32 [B,T], puncts i32 [B,T], mask bool [B,T], lens i32)."""
    lens = np.array([len(s) for s in phone_seqs], dtype=np.int32)
    T = int(lens.max()) if len(lens) else 0
    ph = np.zeros((len(lens), T), dtype=np.int32)
    pu = np.zeros((len(lens), T), dtype=np.int32)
    for b, (s, q) in enumerate(zip(phone_seqs, punct_seqs)):
        ph[b, :len(s)] = s
        pu[b, :len(q)] = q
    mask = np.arange(T)[None, :] >= lens[:, None]
    return ph, pu, mask, lens
