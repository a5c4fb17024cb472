#!/usr/bin/env python
"""Pin oracle/frontend_oracle.py against the reference's own code and write tests/golden/{tokeniser,melfront}.npz.

Runs ONLY in the build container (needs /root/reference; never on the GPU box).

  * tokeniser: imports the reference's real ``zerovox.tts.symbols.Symbols`` and
    ``zerovox.tts.synthesize.ZeroVoxTTS`` (packages that are absent here and that the tokeniser never touches —
    torchinfo, librosa, lightning, the NeMo normaliser — are stubbed as empty modules) and runs the UNBOUND
    ``ZeroVoxTTS.transcript2phonemids`` on a table of transcripts.
  * mel front-end: runs the reference's real ``zerovox.tts.mels.get_mel_from_wav`` (mels.py:356-394: padding,
    np.abs, np.dot, log-clip, energy norm are the reference's own lines) with the two librosa calls it makes served
    by independent third-party implementations that are present: ``librosa.stft`` -> torch.stft (float64, the
    formulation the reference keeps commented out at mels.py:330-343) and ``librosa.filters.mel`` ->
    torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='slaney').  librosa itself is not installable
    offline; see the pinning note in oracle/frontend_oracle.py.

Usage:  python oracle/make_goldens_frontend.py   (from the repo root)
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("ZEROVOX_REFERENCE", "/root/reference")

from oracle import frontend_oracle as fo  # noqa: E402
from zerovox_b200.synthetic import make_speech_like  # noqa: E402

PHONES_EN = "'-abcdefghijklmnopqrstuvwxyz"       # configs/tts_medium.yaml:20
PUNCTS_EN = " ,.;:-!?\""                          # configs/tts_medium.yaml:21
PHONES_UNI = "aäbcdeəfghiɪjklmnŋoöprsʃtuüvzʒß"   # a non-ASCII vocabulary (code points, not bytes)
PUNCTS_UNI = " ,.¿?!…"

TRANSCRIPTS = [
    "this is a test.",
    "entweder zu helfen, wenn",
    "  leading blanks and trailing ones   ",
    "...starts with punctuation, then words!",
    "a,b;c:d-e!f?g\"h",
    "double  blanks   and , mixed ;. punctuation ?!",
    "UPPER case and d1g1ts 42 are skipped",
    "it's a well-known fact",
    "unknown # char between, # , runs",
    "",
    "   ",
    "?",
    "x",
    "ends with a dash -",
]
TRANSCRIPTS_UNI = [
    "ʃöne grüße, ɪŋ…",
    "¿ke tal? muj bjen!",
    "straße… əŋ ʒ",
]


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _librosa_stub():
    import torchaudio

    def stft(y, n_fft, hop_length, win_length, window, center):
        assert window == "hann" and center is False
        w = torch.hann_window(win_length, periodic=True, dtype=torch.float64)
        s = torch.stft(torch.from_numpy(np.asarray(y)).double(), n_fft, hop_length=hop_length, win_length=win_length,
                       window=w, center=False, normalized=False, onesided=True, return_complex=True)
        return s.numpy().astype(np.complex64)        # librosa.stft returns complex64 for float32 input

    def mel(sr, n_fft, n_mels, fmin, fmax):
        fb = torchaudio.functional.melscale_fbanks(n_freqs=1 + n_fft // 2, f_min=float(fmin), f_max=float(fmax),
                                                   n_mels=n_mels, sample_rate=sr, norm="slaney", mel_scale="slaney")
        return fb.T.numpy().astype(np.float32)

    m = _stub("librosa", stft=stft)
    m.filters = _stub("librosa.filters", mel=mel)
    m.effects = _stub("librosa.effects")
    m.util = _stub("librosa.util")
    return m


def main():
    import torch.nn as nn
    _stub("torchinfo", summary=lambda *a, **k: None)
    _stub("lightning", LightningModule=type("LightningModule", (nn.Module,), {}),
          LightningDataModule=type("LightningDataModule", (), {}))
    _librosa_stub()
    _stub("zerovox.tts.normalize", ZeroVoxNormalizer=object)
    for name in ("scipy.io.wavfile",):
        __import__(name)
    sys.path.insert(0, REF)
    from zerovox.tts.symbols import Symbols as RefSymbols
    from zerovox.tts.synthesize import ZeroVoxTTS
    from zerovox.tts import mels as refmels

    out = os.path.join(ROOT, "tests", "golden")

    # ------------------------------------------------------------------ tokeniser
    cases = []
    for phones, puncts, texts in ((PHONES_EN, PUNCTS_EN, TRANSCRIPTS), (PHONES_UNI, PUNCTS_UNI, TRANSCRIPTS_UNI)):
        rs = RefSymbols(phones, puncts)
        os_ = fo.Symbols(phones, puncts)
        assert rs.num_phones == os_.num_phones and rs.num_puncts == os_.num_puncts
        me = types.SimpleNamespace(_symbols=rs)
        for t in texts:
            ph, pu = ZeroVoxTTS.transcript2phonemids(me, t)
            oph, opu = fo.transcript2phonemids(os_, t)
            assert (ph, pu) == (oph, opu), (t, ph, oph, pu, opu)
            cases.append({"phones": phones, "puncts": puncts, "text": t, "phone_ids": ph, "punct_ids": pu})
    with open(os.path.join(out, "tokeniser.json"), "w", encoding="utf-8") as f:
        json.dump(cases, f, ensure_ascii=False, indent=1)
    print(f"tokeniser.json: {len(cases)} transcripts, restatement == reference")

    # ------------------------------------------------------------------ mel front-end
    gold = {}
    for name, (n, seed) in {"short": (4000, 1), "prompt": (22050 * 2 + 123, 2)}.items():
        wav = make_speech_like(n, seed=seed)
        refmels.mel_basis = None
        spec, energy = refmels.get_mel_from_wav(audio=wav, sampling_rate=22050, fft_size=1024, hop_size=256,
                                                win_length=1024, num_mels=80, fmin=0, fmax=8000)
        ospec, oenergy = fo.get_mel_from_wav(wav)
        e1 = float(np.abs(spec - ospec).max())
        e2 = float((np.abs(energy - oenergy) / np.abs(energy).max()).max())
        print(f"melfront[{name}]: frames {spec.shape[1]}  |restatement - reference| log-mel {e1:.2e}  energy rel {e2:.2e}")
        assert spec.shape == ospec.shape and e1 < 2e-4 and e2 < 1e-5
        gold[name + "_n"] = n
        gold[name + "_seed"] = seed
        gold[name + "_spec"] = spec.astype(np.float32)
        gold[name + "_energy"] = energy.astype(np.float32)
    np.savez_compressed(os.path.join(out, "melfront.npz"), **gold)
    print("wrote tests/golden/melfront.npz")


if __name__ == "__main__":
    main()
