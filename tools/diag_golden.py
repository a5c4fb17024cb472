"""Diagnostic (GPU): mel error of policy 0 / 1 against a golden case; ZVX_TC_MASK bisects the tensor-core sites."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import zerovox_oracle as zo
from zerovox_b200.testing import build_model
name = sys.argv[1] if len(sys.argv) > 1 else "tiny_predicted"
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
cfg = zo.ZeroVoxConfig.tiny() if name.startswith("tiny") else zo.ZeroVoxConfig()
w = zo.make_weights(cfg, seed=int(g["seed_w"]), dur_bias=float(g["dur_bias"]))
x = zo.make_inputs(cfg, int(g["B"]), int(g["T"]), int(g["T_ref"]), seed=int(g["seed_x"]), ragged=bool(g["ragged"]),
                   dur_lo=int(g["dur_lo"]), dur_hi=int(g["dur_hi"]))
for pol in (0, 1):
    m = build_model(cfg, w, device="cuda:0", tensor_core_policy=pol)
    with torch.no_grad():
        wav, mel, mel_len, logd = m(dict(x), force_duration=bool(g["force"]))
    e = (mel.cpu() - torch.from_numpy(g["mel"])).abs()
    print(f"mask={os.environ.get('ZVX_TC_MASK')} policy {pol} mel_len {mel_len.tolist()} mel err max {e.max().item():.4f} per-utt {[round(v, 4) for v in e.amax(dim=(1, 2)).tolist()]}")
