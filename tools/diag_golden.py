"""Diagnostic: per-frame error of policy 0 / 1 against a golden case (GPU)."""
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
outs = {}
for pol in (0, 1):
    m = build_model(cfg, w, device="cuda:0", tensor_core_policy=pol)
    with torch.no_grad():
        wav, mel, mel_len, logd = m(dict(x), force_duration=bool(g["force"]))
    outs[pol] = (wav.cpu(), mel.cpu(), mel_len.cpu())
    print("policy", pol, "mel_len", mel_len.tolist(), "golden", g["mel_len"].tolist())
    e = (mel.cpu() - torch.from_numpy(g["mel"])).abs()
    print("  mel err max", e.max().item(), "per-utt max", e.amax(dim=(1, 2)).tolist())
    print("  per-frame max (utt0):", np.round(e[0].amax(dim=0).numpy(), 3).tolist())
    if e.shape[0] > 1:
        print("  per-frame max (utt1):", np.round(e[1].amax(dim=0).numpy(), 3).tolist())
