"""Diagnostic (GPU): signed error statistics of the tcgen05 TF32 GEMM on positive operands — truncation shows up as a
negative bias of ~1e-3 relative, round-to-nearest as ~0."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from zerovox_b200.engine import Engine, EngineConfig
from test_gpu_gemm import run
eng = Engine(EngineConfig(), "cuda:0")
g = torch.Generator().manual_seed(0)
M, N, K = 512, 256, 1024
A = torch.rand(M, K, generator=g) + 0.5
W = (torch.rand(1, N, K, generator=g) + 0.5)
ref = (A.double() @ W[0].double().T)
for tc in (0, 1):
    got = run(eng, 0, A, W, use_tc=tc).double()
    rel = (got - ref) / ref
    print(f"use_tc={tc} env={os.environ.get('ZVX_TMAP_F32')} mean signed rel err {rel.mean().item():+.3e}  rms {rel.pow(2).mean().sqrt().item():.3e}")
# operands pre-rounded to tf32 (RN) on the host: tensor-core result should then be exact to fp32 accumulation
def rn_tf32(t):
    i = t.view(torch.int32)
    r = ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF)
    return r.view(torch.float32)
got = run(eng, 0, rn_tf32(A.clone()), rn_tf32(W.clone()), use_tc=1).double()
rel = (got - ref) / ref
print(f"pre-rounded RN operands: mean signed rel err {rel.mean().item():+.3e} rms {rel.pow(2).mean().sqrt().item():.3e}")
