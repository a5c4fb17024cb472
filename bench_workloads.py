"""bench.py --workload config3 | config5: the BASELINE.json configs that are not the headline line, under the same JSON
contract (value / roofline / cpu_baseline / e2e), single GPU.

config3: hifigan.Generator alone (hifigan.py:114-130), mel length L in {128..4096} x batch B in {1..128}, topologies V1
         (dense 512..32-channel convs: the conv roofline) and V2; headline value = V1 at the largest point.
config5: one 4096-phoneme utterance (L ~ 24.6 k frames) through inference_ex with the chunked length regulator / attention
         and the overlap-discard vocoder (model.py:308-347).
"""
from __future__ import annotations

import json
import os
import statistics
import time

import torch

import bench as B_


def _timed(fn, iters, dev, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize(dev)
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def _class_profile(eng, fn):
    eng.profile(True)
    fn()
    prof = eng.profile_read()
    eng.profile(False)
    return prof


def run_config3(args, dev):
    from zerovox_b200 import synthetic as syn
    from zerovox_b200.testing import build_generator
    peaks = B_.measured_peaks()
    tp = B_.tf32_peak(peaks)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    points, head = [], None
    for v in ("v2", "v1"):
        h = getattr(syn.HifiGanConfig, v)()
        hw = syn.make_hifigan_weights(h, torch.Generator().manual_seed(11))
        gen = build_generator(h, {"_meldec." + k: t for k, t in hw.items()}).to(dev)
        for L in (128, 512, 2048, 4096):
            for B in (1, 8, 32, 128):
                if B * L > (131072 if v == "v2" else 65536):
                    continue
                mel = torch.randn((B, 80, L), generator=torch.Generator().manual_seed(1)).to(dev)
                with torch.no_grad():
                    ms = _timed(lambda: gen(mel), max(3, min(args.steps, 5)), dev, flush)
                frames = B * L
                tf = frames * MF[v] * 1e6 / (ms * 1e-3) / 1e12
                pt = {"vocoder": v, "B": B, "L": L, "ms": round(ms, 4), "mel_frames_per_sec": frames / ms * 1e3,
                      "audio_sec_per_sec": frames * 256 / 22050 / ms * 1e3, "tflops": tf, "frac_of_tf32_peak": tf / tp["sustained"]}
                points.append(pt)
                if v == "v1" and (head is None or frames >= head[0]["B"] * head[0]["L"]):
                    head = (pt, gen, mel)
    pt, gen, mel = head
    eng = gen._ctx.get(dev) if hasattr(gen, "_ctx") else None
    line = {"metric": "audio-sec/sec (RTF^-1) @22.05kHz, mel->waveform (HiFi-GAN Generator only)", "value": pt["audio_sec_per_sec"],
            "unit": B_.UNIT, "n_gpus": 1, "steps": max(3, min(args.steps, 5)), "warmup": 3, "ms_per_step": pt["ms"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32 (tcgen05, fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": f"config 3: HiFi-GAN V1 Generator only, B={pt['B']} x L={pt['L']} mel frames (headline point of the sweep)",
                       "l2": "256 MiB buffer written between timed runs"},
            "points": points}
    if eng is not None:
        with torch.no_grad():
            prof = _class_profile(eng, lambda: gen(mel))
        dom = max(prof, key=lambda k: prof[k]["ms"])
        d = prof[dom]
        ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": tp["sustained"], "unit": "TFLOP/s",
                            "frac": ach / tp["sustained"], "peak_source": "TF32 dense, sustained; " + tp["source"],
                            "traffic": None, "launches_per_step": d["launches"], "avg_launch_ms": d["ms"] / d["launches"],
                            "share_of_step": d["ms"] / pt["ms"],
                            "whole_generator_tflops": pt["tflops"], "whole_generator_frac": pt["frac_of_tf32_peak"]}
        line["gpu_launches"] = int(sum(v["launches"] for v in prof.values()))
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = _cpu_generator(syn)
    print(json.dumps(line), flush=True)


MF = B_.MFLOP_PER_FRAME


def _cpu_generator(syn):
    """The reference's hifigan.Generator (oracle/_ref) on the host cores at bounded points of the sweep."""
    from oracle import reference_modules as rm
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    out = {"unit": B_.UNIT, "cores": threads, "points": []}
    if rm.available():
        model, _, hifigan = rm.import_reference()
        out["kind"] = "reference"
    else:
        out["kind"] = "port"
    for v, Bc, Lc in (("v2", 8, 512), ("v1", 2, 256)):
        h = getattr(syn.HifiGanConfig, v)()
        hw = syn.make_hifigan_weights(h, torch.Generator().manual_seed(11))
        mel = torch.randn((Bc, 80, Lc), generator=torch.Generator().manual_seed(1))
        if out["kind"] == "reference":
            gen = hifigan.Generator(model.AttrDict(h.as_json_dict())).eval()
            gen.remove_weight_norm()
            gen.load_state_dict(hw)
            fn = lambda: gen(mel)   # noqa: E731
        else:
            from oracle import zerovox_oracle as zo
            fn = lambda: zo.hifigan_generator(h, hw, mel, prefix="")   # noqa: E731
        with torch.no_grad():
            fn()
            t0 = time.perf_counter()
            fn()
            t = time.perf_counter() - t0
        out["points"].append({"vocoder": v, "B": Bc, "L": Lc, "seconds": t, "audio_sec_per_sec": Bc * Lc * 256 / 22050 / t,
                              "gflops": Bc * Lc * MF[v] * 1e6 / t / 1e9})
    out["value"] = out["points"][-1]["audio_sec_per_sec"]
    out["sample"] = "V1 B=2 x L=256 (value) and V2 B=8 x L=512, one run after one warm-up, all host threads"
    return out


def run_config5(args, dev):
    from zerovox_b200 import synthetic as syn
    from zerovox_b200.testing import build_model
    cfg = syn.ZeroVoxConfig()
    w = syn.make_weights(cfg, seed=0)
    model = build_model(cfg, w, device=dev, tensor_core_policy=args.policy)
    eng = model._shared_ctx.get(dev)
    T, chunk = 4096, 2048
    x = syn.make_inputs(cfg, 1, T, 440, seed=11)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = B_.measured_peaks()
    with torch.no_grad():
        style = model._spkemb(x["ref_mel"].to(dev))
        x1 = {k: v.to(dev) for k, v in x.items() if k != "ref_mel"}
        out = {}

        def run():
            out["r"] = model.inference_ex(x1, style_embed=style, force_duration=True, vocoder_chunk_frames=chunk)
        n = max(3, min(args.steps, 5))
        ms = _timed(run, n, dev, flush)
        mel_len = out["r"][1]
        # e2e: host ids / durations in, waveform out to pinned host memory
        xh = {k: v.pin_memory() for k, v in x.items() if k != "ref_mel"}
        wav_host = torch.empty((mel_len * cfg.hop_length,), dtype=torch.float32).pin_memory()

        def run_e2e():
            r = model.inference_ex(xh, style_embed=style, force_duration=True, vocoder_chunk_frames=chunk)
            wav_host.copy_(r[0], non_blocking=True)
            torch.cuda.synchronize(dev)
        for _ in range(2):
            run_e2e()
        t0 = time.perf_counter()
        for _ in range(n):
            run_e2e()
        t_e2e = (time.perf_counter() - t0) / n
        l0 = eng.launch_count()
        run()
        launches = eng.launch_count() - l0
        rf = B_.roofline_block(eng, run, peaks, ms, None)
    audio = mel_len * cfg.hop_length / cfg.sampling_rate
    line = {"metric": B_.METRIC, "value": audio / (ms * 1e-3), "unit": B_.UNIT, "n_gpus": 1, "steps": n, "warmup": 3,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 (tcgen05, fp32 accumulate); 3xTF32 split encoder", "data": "synthetic",
            "config": {"workload": f"config 5: one {T}-phoneme utterance, forced durations U{{2..10}} -> {mel_len} mel frames "
                                   f"({audio:.1f} s of audio); query-chunked attention, vocoder in chunks of {chunk} frames with a "
                                   "14-frame discarded halo", "mel_frames_per_sec": mel_len / (ms * 1e-3),
                       "l2": "256 MiB buffer written between timed runs"},
            "gpu_launches": launches,
            "e2e": {"value": audio / t_e2e, "unit": B_.UNIT, "ms_per_step": t_e2e * 1e3,
                    "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in xh.values()),
                    "d2h_bytes_per_step": wav_host.numel() * 4}}
    if rf:
        line["roofline"] = rf
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = _cpu_longform(syn, cfg, w)
    print(json.dumps(line), flush=True)


def _cpu_longform(syn, cfg, w, T=1024):
    """the reference's inference_ex at T = 1024 phonemes (T = 4096 materialises several ~4.8 GB
    attention temporaries per layer on the CPU), all host threads, one run."""
    from oracle import reference_modules as rm
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    x = syn.make_inputs(cfg, 1, T, 440, seed=11)
    x1 = {k: v for k, v in x.items() if k != "ref_mel"}
    if rm.available():
        zv = rm.build_reference_model(cfg, w)
        kind = "reference"
        with torch.no_grad():
            style = zv._spkemb(x["ref_mel"])
            t0 = time.perf_counter()
            r = zv.inference_ex(dict(x1), style_embed=style, force_duration=True)
            t = time.perf_counter() - t0
        mel_len = int(r[1])
    else:
        from oracle import zerovox_oracle as zo
        kind = "port"
        with torch.no_grad():
            style = zo.speaker_embed(cfg, w, x["ref_mel"])
            t0 = time.perf_counter()
            r = zo.zerovox_inference_ex(cfg, w, dict(x1), style, force_duration=True)
            t = time.perf_counter() - t0
        mel_len = int(r[1])
    return {"value": mel_len * cfg.hop_length / cfg.sampling_rate / t, "unit": B_.UNIT, "cores": threads, "kind": kind,
            "sample": f"one {T}-phoneme utterance ({mel_len} frames) through inference_ex, one run, all host threads",
            "seconds_per_run": t}


def run(args, rank, local_rank, world):
    if rank != 0:
        return                                         # these workloads do not shard: one GPU
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if args.workload == "config3":
        run_config3(args, dev)
    else:
        run_config5(args, dev)
