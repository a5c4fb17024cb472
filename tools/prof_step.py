"""One configs[1] step inside a cudaProfilerStart/Stop range, for ncu (--profile-from-start off).

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:voc -o gpurun_out/x \
        python tools/prof_step.py [--batch 32] [--phonemes 128] [--policy 1] [--stage all|vocoder]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from zerovox_b200 import synthetic as syn  # noqa: E402
from zerovox_b200.testing import build_model  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--batch", type=int, default=32)
    p.add_argument("--phonemes", type=int, default=128)
    p.add_argument("--ref-frames", type=int, default=440)
    p.add_argument("--policy", type=int, default=1)
    p.add_argument("--warmup", type=int, default=2)
    args = p.parse_args()
    dev = torch.device("cuda", 0)
    cfg = syn.ZeroVoxConfig()
    w = syn.make_weights(cfg, seed=0)
    x = {k: v.to(dev) for k, v in syn.make_inputs(cfg, args.batch, args.phonemes, args.ref_frames, seed=7).items()}
    model = build_model(cfg, w, device=dev, tensor_core_policy=args.policy)
    with torch.no_grad():
        for _ in range(args.warmup):
            model(x, force_duration=True)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        model(x, force_duration=True)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    print("prof_step done")


if __name__ == "__main__":
    main()
