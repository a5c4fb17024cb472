"""Tiny-config forward (every kernel class) for compute-sanitizer runs:
   compute-sanitizer --tool memcheck python tools/sanity_tiny.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from zerovox_b200 import synthetic as syn  # noqa: E402
from zerovox_b200.testing import build_model  # noqa: E402

policy = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = syn.ZeroVoxConfig.tiny()
w = syn.make_weights(cfg, seed=1)
model = build_model(cfg, w, device="cuda:0", tensor_core_policy=policy)
for force in (True, False):
    x = syn.make_inputs(cfg, 3, 11, 24, seed=7, ragged=True, dur_lo=0, dur_hi=5)
    with torch.no_grad():
        wav, mel, mel_len, logd = model(x, force_duration=force)
    torch.cuda.synchronize()
    print("force" if force else "predicted", "mel_len", mel_len.tolist(), "wav", tuple(wav.shape),
          "finite", bool(torch.isfinite(wav).all()))
print("sanity ok")
