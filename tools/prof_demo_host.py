"""Host-side profile of the single-utterance path (ZeroVoxTTS.tts): cProfile over N iterations + CUDA-event time of the same
loop, to split wall time into GPU work and host overhead."""
import cProfile
import dataclasses
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from zerovox_b200 import synthetic as syn  # noqa: E402
from zerovox_b200.testing import build_model  # noqa: E402
from zerovox_b200.tts.symbols import Symbols  # noqa: E402
from zerovox_b200.tts.synthesize import ZeroVoxTTS  # noqa: E402


class N:
    def normalize(self, t):
        return t.lower(), None


cfg = dataclasses.replace(syn.ZeroVoxConfig(), decoder_kind="styletts")
w = syn.make_weights(cfg, seed=0, dur_bias=float(np.log(7.0)))
model = build_model(cfg, w, device="cuda:0")
tts = ZeroVoxTTS(language="en", syms=Symbols(cfg.phones, cfg.puncts), checkpoint=None, meldec_model=None, hop_length=256,
                 sampling_rate=22050, n_mel_channels=80, fft_size=1024, win_length=1024, mel_fmin=0, mel_fmax=8000,
                 infer_device="cuda:0", model=model, normalizer=N())
spk = tts.speaker_embed(syn.make_speech_like(5 * 22050, seed=1))
text = "this is a test of the zerovox engine, on a b two hundred."
for _ in range(5):
    tts.tts(text, spk)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
for _ in range(50):
    tts.tts(text, spk)
b.record()
torch.cuda.synchronize()
print(f"wall {1e3 * (time.perf_counter() - t0) / 50:.3f} ms / call; CUDA-event span {a.elapsed_time(b) / 50:.3f} ms / call")
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    tts.tts(text, spk)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
