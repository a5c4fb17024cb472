#!/usr/bin/env python
"""Measurements for the BASELINE.json configs that are not the bench.py headline (configs[1]):

    python tools/bench_configs.py --config 1 --cpu      # one utterance, styledec model, GPU latency vs CPU oracle
    python tools/bench_configs.py --config 3            # HiFi-GAN Generator only: mel length x batch sweep (V1, V2)
    python tools/bench_configs.py --config 5            # long-form: one 4096-phoneme utterance, chunked vocoder
    torchrun --nproc-per-node N tools/bench_configs.py --config 4   # B=256 ragged batch sharded over N GPUs (NCCL)
    python tools/bench_configs.py --config 6 --cpu      # speaker-prompt front-end (trim + log-mel) and tokeniser / collator

CUDA-event timing on the device, 3 warm-ups, median of --iters runs; one JSON object per measurement on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from zerovox_b200 import synthetic as syn  # noqa: E402
from zerovox_b200.testing import build_generator, build_model  # noqa: E402

MFLOP_PER_FRAME = {"v1": 614.1, "v2": 38.5, "v3": 45.0}


def timed(fn, iters, dev):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize(dev)
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def config1(args, dev):
    """configs[0]: one utterance through inference_ex, tts_medium_styledec + HiFi-GAN V2 (the shipped default model shape),
    precomputed speaker embedding as in zerovox/demo.py; GPU latency next to the CPU oracle on this box's cores."""
    import dataclasses
    import time
    cfg = dataclasses.replace(syn.ZeroVoxConfig(), decoder_kind="styletts")
    w = syn.make_weights(cfg, seed=0)
    model = build_model(cfg, w, device=dev)
    x = syn.make_inputs(cfg, 1, 128, 440, seed=7)
    with torch.no_grad():
        style = model._spkemb(x["ref_mel"].to(dev))
        x1 = {k: v.to(dev) for k, v in x.items() if k != "ref_mel"}
        out = {}

        def run():
            out["r"] = model.inference_ex(x1, style_embed=style, force_duration=True)
        ms = timed(run, max(args.iters, 20), dev)
        t0 = time.perf_counter()
        for _ in range(20):
            run()
        torch.cuda.synchronize(dev)
        wall_ms = (time.perf_counter() - t0) / 20 * 1e3
    mel_len = out["r"][1]
    rec = {"config": 1, "decoder": "styletts", "vocoder": "v2", "B": 1, "phonemes": 128, "mel_frames": mel_len,
           "gpu_ms_events": round(ms, 3), "gpu_ms_wall_incl_launch": round(wall_ms, 3), "audio_sec": mel_len * 256 / 22050,
           "rtf_inverse_gpu": mel_len * 256 / 22050 / (wall_ms * 1e-3)}
    if args.cpu:
        from oracle import zerovox_oracle as zo   # measurement tool: CPU baseline leg
        torch.set_num_threads(os.cpu_count() or 1)
        xs = {k: v for k, v in x.items() if k != "ref_mel"}
        with torch.no_grad():
            zo.zerovox_inference_ex(cfg, w, dict(xs), style.cpu(), force_duration=True)
            t0 = time.perf_counter()
            for _ in range(3):
                zo.zerovox_inference_ex(cfg, w, dict(xs), style.cpu(), force_duration=True)
            cpu_s = (time.perf_counter() - t0) / 3
        rec.update({"cpu_ms": round(cpu_s * 1e3, 1), "cpu_threads": os.cpu_count(), "rtf_inverse_cpu": mel_len * 256 / 22050 / cpu_s})
    print(json.dumps(rec), flush=True)


def config2(args, dev):
    """configs[1] shape (B = 32 x T = 128, forced durations) for either decoder kind; stage split from CUDA events."""
    import dataclasses
    cfg = dataclasses.replace(syn.ZeroVoxConfig(), decoder_kind=args.decoder)
    w = syn.make_weights(cfg, seed=0)
    model = build_model(cfg, w, device=dev)
    eng = model._shared_ctx.get(dev)
    x = {k: v.to(dev) for k, v in syn.make_inputs(cfg, 32, 128, 440, seed=7).items()}
    with torch.no_grad():
        out = {}

        def run():
            out["r"] = model(x, force_duration=True)
        ms = timed(run, args.iters, dev)
        style = eng.spkemb(x["ref_mel"])
        r = eng.encode(x["phoneme"], x["puncts"], style, None, x["duration"])
        feats = eng.length_regulate(r["xprime"], r["duration_rounded"], r["L_max"])
        mel = eng.decode(feats, style, mel_len=r["mel_len"], want_blc=False)[1]
        stages = {"spkemb": timed(lambda: eng.spkemb(x["ref_mel"]), args.iters, dev),
                  "encode": timed(lambda: eng.encode(x["phoneme"], x["puncts"], style, None, x["duration"]), args.iters, dev),
                  "decode": timed(lambda: eng.decode(feats, style, mel_len=r["mel_len"], want_blc=False), args.iters, dev),
                  "vocode": timed(lambda: eng.vocode(mel), args.iters, dev)}
    frames = int(out["r"][2].sum())
    print(json.dumps({"config": 2, "decoder": args.decoder, "ms": round(ms, 3), "stage_ms": {k: round(v, 3) for k, v in stages.items()},
                      "mel_frames": frames, "audio_sec_per_sec": frames * 256 / 22050 / ms * 1e3}), flush=True)


def config3(args, dev):
    for v in ("v2", "v1"):
        h = getattr(syn.HifiGanConfig, v)()
        hw = syn.make_hifigan_weights(h, torch.Generator().manual_seed(11))
        gen = build_generator(h, {"_meldec." + k: t for k, t in hw.items()}).to(dev)
        for L in (128, 512, 2048, 4096):
            for B in (1, 8, 32, 128):
                if B * L > (131072 if v == "v2" else 32768):
                    continue
                mel = torch.randn((B, 80, L), generator=torch.Generator().manual_seed(1)).to(dev)
                with torch.no_grad():
                    ms = timed(lambda: gen(mel), args.iters, dev)
                frames = B * L
                print(json.dumps({"config": 3, "vocoder": v, "B": B, "L": L, "ms": round(ms, 4),
                                  "mel_frames_per_sec": frames / ms * 1e3, "audio_sec_per_sec": frames * 256 / 22050 / ms * 1e3,
                                  "tflops": frames * MFLOP_PER_FRAME[v] * 1e6 / (ms * 1e-3) / 1e12}), flush=True)


def config5(args, dev):
    cfg = syn.ZeroVoxConfig()
    w = syn.make_weights(cfg, seed=0)
    model = build_model(cfg, w, device=dev)
    x = syn.make_inputs(cfg, 1, 4096, 440, seed=11)
    with torch.no_grad():
        style = model._spkemb(x["ref_mel"].to(dev))
        x1 = {k: v.to(dev) for k, v in x.items() if k != "ref_mel"}
        out = {}

        def run():
            out["r"] = model.inference_ex(x1, style_embed=style, force_duration=True, vocoder_chunk_frames=args.chunk)
        ms = timed(run, args.iters, dev)
    mel_len = out["r"][1]
    print(json.dumps({"config": 5, "phonemes": 4096, "mel_frames": mel_len, "vocoder_chunk_frames": args.chunk, "ms": round(ms, 3),
                      "audio_sec": mel_len * 256 / 22050, "audio_sec_per_sec": mel_len * 256 / 22050 / ms * 1e3,
                      "mel_frames_per_sec": mel_len / ms * 1e3}), flush=True)


def config4(args, dev, rank, world):
    import torch.distributed as dist
    from zerovox_b200.parallel import sharded_forward
    cfg = syn.ZeroVoxConfig()
    w = syn.make_weights(cfg, seed=0)
    model = build_model(cfg, w, device=dev)
    x = None
    if rank == 0:
        g = torch.Generator().manual_seed(3)
        B, T = args.batch, 192
        x = syn.make_inputs(cfg, B, T, 440, seed=13)
        lens = torch.randint(64, 193, (B,), generator=g)     # T_i ~ U{64..192}, padded to 192 with phoneme_mask
        mask = torch.arange(T)[None, :] >= lens[:, None]
        x["phoneme_mask"] = mask
        for k in ("phoneme", "puncts", "duration"):
            x[k] = x[k].masked_fill(mask, 0)
    out = {}

    def run():
        with torch.no_grad():
            out["r"] = sharded_forward(model, x, force_duration=True, device=dev)
    for _ in range(2):
        run()
    ts = []
    for _ in range(args.iters):
        dist.barrier()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t))
    if rank == 0:
        ms = statistics.median(ts)
        wav, mel, mel_len, _ = out["r"]
        frames = int(mel_len.sum())
        print(json.dumps({"config": 4, "n_gpus": world, "B": args.batch, "ms_incl_scatter_gather": round(ms, 3),
                          "mel_frames": frames, "audio_sec_per_sec": frames * 256 / 22050 / ms * 1e3,
                          "mel_frames_per_sec": frames / ms * 1e3, "wav_shape": list(wav.shape)}), flush=True)


def config6(args, dev):
    """SURVEY.md 8f rows 3 + 4: the speaker-prompt front-end (trim + log-mel, csrc/frontend.cu) on B prompts of 5 s, and the
    native tokeniser / collator; CPU oracle beside both with --cpu (bounded samples)."""
    import time
    import numpy as np
    from zerovox_b200.frontend import MelFrontend, Tokeniser
    fe = MelFrontend(device=dev)
    n = 22050 * 5
    for B in (1, 32, 256):
        wavs = torch.from_numpy(np.stack([syn.make_speech_like(n, seed=s) for s in range(min(B, 8))])).repeat((B + 7) // 8, 1)[:B]
        wavs = wavs.contiguous().to(dev)
        F = fe.num_frames(n)
        ms_mel = timed(lambda: fe.mel(wavs, with_energy=True), max(args.iters, 20), dev)
        ms_all = timed(lambda: fe.speaker_prompt_mel(wavs), max(args.iters, 20), dev)
        alg_bytes = B * F * (256 * 4 + 80 * 4 + 4)           # one hop of samples in, one mel row + energy out
        rec = {"config": 6, "what": "mel front-end", "B": B, "samples": n, "frames": B * F, "mel_ms": round(ms_mel, 4),
               "trim_plus_mel_ms_incl_sync": round(ms_all, 4), "frames_per_sec": B * F / ms_mel * 1e3,
               "audio_sec_per_sec": B * n / 22050 / ms_mel * 1e3, "algorithmic_GBps": alg_bytes / ms_mel / 1e6,
               "fft_gflops": B * F * 5 * 1024 * 10 / ms_mel / 1e6}
        if args.cpu and B == 32:
            from oracle import frontend_oracle as fo      # measurement tool: CPU baseline leg
            w1 = wavs[0].cpu().numpy()
            fo.speaker_prompt_mel(w1)
            t0 = time.perf_counter()
            for _ in range(5):
                fo.speaker_prompt_mel(w1)
            cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
            rec.update({"cpu_ms_per_prompt_numpy_oracle": round(cpu_ms, 2), "cpu_audio_sec_per_sec": 5.0 / cpu_ms * 1e3})
        print(json.dumps(rec), flush=True)
    import random
    rng = random.Random(1)
    phones, puncts = "'-abcdefghijklmnopqrstuvwxyz", " ,.;:-!?\""
    tok = Tokeniser(phones, puncts)
    texts = ["".join(rng.choice(phones + puncts + "   ") for _ in range(160)) for _ in range(2048)]
    t0 = time.perf_counter()
    ids = [tok.transcript2phonemids(t) for t in texts]
    t1 = time.perf_counter()
    tok.collate([i[0] for i in ids], [i[1] for i in ids], pinned=True)
    t2 = time.perf_counter()
    rec = {"config": 6, "what": "tokeniser + collator (host C via ctypes)", "transcripts": len(texts), "chars": 160 * len(texts),
           "tokenise_us_per_transcript": (t1 - t0) / len(texts) * 1e6, "collate_ms_batch_2048": (t2 - t1) * 1e3}
    if args.cpu:
        from oracle import frontend_oracle as fo
        sym = fo.Symbols(phones, puncts)
        t0 = time.perf_counter()
        for t in texts:
            fo.transcript2phonemids(sym, t)
        rec["python_oracle_us_per_transcript"] = (time.perf_counter() - t0) / len(texts) * 1e6
    print(json.dumps(rec), flush=True)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--config", type=int, required=True, choices=[1, 2, 3, 4, 5, 6])
    p.add_argument("--decoder", default="fastspeech2", choices=["fastspeech2", "styletts"])
    p.add_argument("--cpu", action="store_true", help="config 1: also time the CPU oracle")
    p.add_argument("--iters", type=int, default=5)
    p.add_argument("--chunk", type=int, default=2048)
    p.add_argument("--batch", type=int, default=256)
    args = p.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if args.config == 1:
        config1(args, dev)
    elif args.config == 2:
        config2(args, dev)
    elif args.config == 3:
        config3(args, dev)
    elif args.config == 5:
        config5(args, dev)
    elif args.config == 6:
        config6(args, dev)
    else:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        config4(args, dev, rank, world)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
