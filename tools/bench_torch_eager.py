#!/usr/bin/env python
"""Context number (not the judged baseline): the same eval forward through plain PyTorch ops on the SAME GPU.

The reference ships no CUDA code of its own — on a GPU it runs PyTorch-eager cuDNN / cuBLAS kernels.
/root/reference does not exist on the GPU box, so this runs the oracle's functional restatement of the reference modules
(oracle/zerovox_oracle.py, pinned to the reference by the goldens) with every tensor on cuda: the same ATen calls the
reference would make, including its per-phoneme `.item()` host syncs in the length regulator.

    python tools/bench_torch_eager.py [--batch 32] [--tf32]
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import zerovox_oracle as zo  # noqa: E402  (measurement tool, not the product)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--batch", type=int, default=32)
    p.add_argument("--phonemes", type=int, default=128)
    p.add_argument("--iters", type=int, default=5)
    p.add_argument("--tf32", action="store_true", help="allow TF32 in cuBLAS / cuDNN (PyTorch default is fp32 matmul, TF32 conv)")
    args = p.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    torch.backends.cudnn.allow_tf32 = bool(args.tf32)
    cfg = zo.ZeroVoxConfig()
    w = {k: v.to(dev) for k, v in zo.make_weights(cfg, seed=0).items()}
    x = {k: v.to(dev) for k, v in zo.make_inputs(cfg, args.batch, args.phonemes, 440, seed=7).items()}
    times, frames = [], 0
    with torch.no_grad():
        for i in range(2 + args.iters):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            wav, mel, mel_len, logd, _ = zo.zerovox_forward(cfg, w, dict(x), force_duration=True)
            torch.cuda.synchronize()
            if i >= 2:
                times.append(time.perf_counter() - t0)
            frames = int(mel_len.sum())
    t = statistics.median(times)
    print(json.dumps({"impl": "pytorch-eager on the same GPU (oracle restatement of the reference modules, cuda tensors)",
                      "tf32": bool(args.tf32), "batch": args.batch, "ms_per_step": t * 1e3,
                      "audio_sec_per_sec": frames * cfg.hop_length / cfg.sampling_rate / t, "mel_frames_per_sec": frames / t}))


if __name__ == "__main__":
    main()
