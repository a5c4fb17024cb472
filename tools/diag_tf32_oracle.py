"""CPU diagnostic: how much does TF32 operand rounding (round-to-nearest, fp32 accumulate) move the oracle's decoder
output on a golden case?  Separates tensor-core precision from kernel bugs."""
import os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import zerovox_oracle as zo

def rn(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)

name = sys.argv[1] if len(sys.argv) > 1 else "tiny_predicted"
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
cfg = zo.ZeroVoxConfig.tiny() if name.startswith("tiny") else zo.ZeroVoxConfig()
w = zo.make_weights(cfg, seed=int(g["seed_w"]), dur_bias=float(g["dur_bias"]))
x = zo.make_inputs(cfg, int(g["B"]), int(g["T"]), int(g["T_ref"]), seed=int(g["seed_x"]), ragged=bool(g["ragged"]),
                   dur_lo=int(g["dur_lo"]), dur_hi=int(g["dur_hi"]))
with torch.no_grad():
    style = zo.speaker_embed(cfg, w, x["ref_mel"])
    enc = zo.fs2_encoder(cfg, w, dict(x), style, force_duration=bool(g["force"]))
    feats, mel_len = enc["features"], enc["mel_len"]
    L = feats.shape[1]
    mask = torch.arange(L)[None, :] >= mel_len[:, None]
    ref = zo.fs2_decoder(cfg, w, feats, mask, style)
    lin, c1d, bmm = F.linear, F.conv1d, torch.bmm
    F.linear = lambda a, b, bias=None: lin(rn(a), rn(b), bias)
    F.conv1d = lambda a, b, bias=None, **kw: c1d(rn(a), rn(b), bias, **kw)
    torch.bmm = lambda a, b: bmm(rn(a), rn(b))
    emu = zo.fs2_decoder(cfg, w, feats, mask, style)
e = (emu - ref).abs()
valid = ~mask
print(name, "decoder mel: max|ref|", ref.abs().max().item(), "TF32-emulated max err", e.max().item(),
      "per-utt", [e[b][valid[b]].max().item() for b in range(e.shape[0])], "rel-rms", (e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item())
