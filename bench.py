#!/usr/bin/env python
"""bench.py — audio-seconds/second (RTF^-1) of the ZeroVOX phoneme -> waveform forward on B200.

    python bench.py --gpus N --steps K --warmup W            # this engine (default)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]): tts_medium (FS2/SCLN decoder) + HiFi-GAN V2, batch = 32 EN phoneme sequences of
length 128, forced durations U{2..10} (L ~ 768 mel frames per utterance), ref_mel [32, 440, 80]; synthetic seeded
weights and inputs (no network for checkpoints).  One "step" = one ZeroVox.forward over one batch: speaker net ->
encoder + variance adaptor -> length regulator -> decoder -> vocoder.  N > 1: one process per GPU, every rank runs
its own batch of 32 (weak scaling, utterances are independent; no data-path collective).

The JSON line follows the driver contract; `value` is device-resident throughput (inputs already in HBM, CUDA-event
time), `e2e` is the same metric through ZeroVox.forward with pinned HOST inputs and the waveform read back to the
host inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "audio-sec/sec (RTF^-1) @22.05kHz, phoneme->waveform"
UNIT = "audio-s/s"
WORKLOAD = "configs[1]: tts_medium + HiFi-GAN V2, B=32 x T=128 phonemes, forced durations U{2..10}, T_ref=440"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=32)
    p.add_argument("--phonemes", type=int, default=128)
    p.add_argument("--ref-frames", type=int, default=440)
    p.add_argument("--policy", type=int, default=1, help="0 = all fp32 FMA, 1 = TF32 tensor cores where allowed")
    p.add_argument("--cpu-sample-batch", type=int, default=4, help="utterances in the bounded CPU sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        # median over the upper half = samples taken under load
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# engine profiling class -> kernel names in profiles/*_ncu_summary.json (tools/ncu_summary.py)
NCU_KERNELS = {"gemm_tf32_tcgen05": ("gemm_tc_kernel<0, 0>", "gemm_tc_kernel<1, 0>"),
               "gemm_3xtf32_tcgen05": ("gemm_tc_kernel<1, 1>",),
               "vocoder_pair_tcgen05": ("voc_poly_kernel<32>", "voc_poly_kernel<16>", "voc_poly_kernel<8>")}


def ncu_traffic(cls):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the class's kernels from the committed
    `ncu --set full` capture of tools/prof_step.py (same workload); None when there is no capture."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.json")))
    if not paths or cls not in NCU_KERNELS:
        return None, None
    with open(paths[-1]) as f:
        summ = json.load(f)
    tot, n = 0.0, 0
    for k in NCU_KERNELS[cls]:
        e = summ.get(k)
        if e and "avg_dram_traffic_bytes" in e:
            tot += e["avg_dram_traffic_bytes"] * e["launches_captured"]
            n += e["launches_captured"]
    if n == 0:
        return None, None
    return tot / n, {"file": os.path.relpath(paths[-1], ROOT), "launches_captured": n}


def audio_seconds(mel_len_total, cfg):
    return mel_len_total * cfg.hop_length / cfg.sampling_rate


def cpu_reference_run(cfg, w, x, threads, runs):
    """The reference algorithm (oracle port of the reference's PyTorch modules) on the host cores."""
    from oracle import zerovox_oracle as zo  # the CPU-baseline leg is the one place bench.py may use the oracle
    torch.set_num_threads(threads)
    times, frames = [], 0
    with torch.no_grad():
        zo.zerovox_forward(cfg, w, dict(x), force_duration=True)  # warm-up
        for _ in range(runs):
            t0 = time.perf_counter()
            _, _, mel_len, _, _ = zo.zerovox_forward(cfg, w, dict(x), force_duration=True)
            times.append(time.perf_counter() - t0)
            frames = int(mel_len.sum())
    return statistics.median(times), frames


def pipeline_roofline(prof, peaks, step_ms, sm_mhz):
    """SURVEY.md 8d pipeline figure: T_roof = sum over kernel classes of max(FLOP / peak_class, bytes / HBM bandwidth),
    with the ALGORITHMIC flops / bytes the engine recorded per class.  peak_class: TF32 tcgen05 = half the measured
    sustained bf16 rate; 3xTF32 split = a third of that (three MMAs per product); fp32 FMA = 148 SMs x 128 lanes x 2 x
    the SM clock sampled during the timed region.  Kernel classes without a flop model (norms, softmax, gathers:
    ~9 % of the step) are outside both sums."""
    tf32 = peaks["bf16_tflops_sustained"] / 2 * 1e12
    fma = 148 * 128 * 2 * sm_mhz * 1e6
    peak_of = {"gemm_tf32_tcgen05": tf32, "vocoder_pair_tcgen05": tf32, "gemm_3xtf32_tcgen05": tf32 / 3,
               "gemm_fp32": fma, "vocoder_conv1d": fma, "vocoder_upsample": fma}
    bw = peaks["hbm_gbs"] * 1e9
    t_roof = t_meas = 0.0
    for k, v in prof.items():
        if v["launches"] <= 0:
            continue
        t_roof += max(v["flops"] / peak_of.get(k, fma), v["bytes"] / bw) * 1e3
        t_meas += v["ms"]
    return {"t_roof_ms": t_roof, "t_measured_ms_modelled_classes": t_meas, "frac": (t_roof / t_meas) if t_meas else None,
            "share_of_step_modelled": t_meas / step_ms, "tf32_peak_tflops": tf32 / 1e12, "fp32_fma_peak_tflops": fma / 1e12}


def run_reference(args, cfg, w, x_full, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nb = min(args.cpu_sample_batch, args.batch)
    x = {k: v[:nb] for k, v in x_full.items()}
    times, frames = [], 0
    from oracle import zerovox_oracle as zo
    torch.set_num_threads(threads)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 1))):
            zo.zerovox_forward(cfg, w, dict(x), force_duration=True)
        for _ in range(args.steps):
            t0 = time.perf_counter()
            _, _, mel_len, _, _ = zo.zerovox_forward(cfg, w, dict(x), force_duration=True)
            times.append(time.perf_counter() - t0)
            frames = int(mel_len.sum())
    t = sum(times) / len(times)
    val = audio_seconds(frames, cfg) / t
    sample = f"{nb} of the {args.batch} utterances of the workload batch per step (same seeds), all host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "mel_frames_per_sec": frames / t,
                   "note": "reference = oracle port of the reference's PyTorch CPU path (pure-Python reference, "
                           "nothing to compile); batched eval tail composed as in oracle/zerovox_oracle.py"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from zerovox_b200 import synthetic as syn
    cfg = syn.ZeroVoxConfig()

    if args.impl == "reference":
        if rank == 0:
            w = syn.make_weights(cfg, seed=0)
            x = syn.make_inputs(cfg, args.batch, args.phonemes, args.ref_frames, seed=7)
            run_reference(args, cfg, w, x, rank)
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU path)"
    from zerovox_b200.testing import build_model
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    w = syn.make_weights(cfg, seed=0)
    x_host = syn.make_inputs(cfg, args.batch, args.phonemes, args.ref_frames, seed=7 + rank)
    x_host = {k: v.pin_memory() for k, v in x_host.items()}
    model = build_model(cfg, w, device=dev, tensor_core_policy=args.policy)
    eng = model._shared_ctx.get(dev)
    x_dev = {k: v.to(dev) for k, v in x_host.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device():
        with torch.no_grad():
            return model(x_dev, force_duration=True)

    for _ in range(max(args.warmup, 3)):
        out = step_device()
    torch.cuda.synchronize(dev)
    mel_len = out[2]
    frames = int(mel_len.sum())
    L_max = int(mel_len.max())

    # ---- timed region: device-resident inputs, CUDA events per step, L2 flushed between steps ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if dist:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = eng.launch_count()
    evs = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_device()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize(dev)
    if dist:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = eng.launch_count() - launches0
    ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: pinned host inputs -> H2D -> forward -> waveform D2H, all inside the timed region ---------------
    e2e = None
    if not args.no_e2e:
        wav_host = torch.empty((args.batch, L_max * cfg.hop_length), dtype=torch.float32).pin_memory()
        len_host = torch.empty((args.batch,), dtype=torch.int64).pin_memory()
        keys = ("phoneme", "puncts", "duration", "ref_mel")
        h2d = sum(x_host[k].numel() * x_host[k].element_size() for k in keys)
        d2h = wav_host.numel() * 4 + len_host.numel() * 8

        def step_e2e():
            with torch.no_grad():
                wav, _, ml, _ = model({k: x_host[k] for k in keys}, force_duration=True)  # forward() does the H2D
                wav_host.copy_(wav, non_blocking=True)
                len_host.copy_(ml, non_blocking=True)
            torch.cuda.synchronize(dev)
        for _ in range(2):
            step_e2e()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        t_e2e = (time.perf_counter() - t0) / args.steps
        e2e = {"ms": t_e2e * 1e3, "h2d": h2d, "d2h": d2h}

    # ---- max over ranks -------------------------------------------------------------------------------------
    if dist:
        t = torch.tensor([ms, e2e["ms"] if e2e else 0.0, float(frames)], device=dev, dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, frames_total = float(tmax[0]), int(tsum[2])
        if e2e:
            e2e["ms"] = float(tmax[1])
    else:
        frames_total = frames

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    value = audio_seconds(frames_total, cfg) / (ms / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32 (tcgen05, fp32 accumulate) decoder/vocoder/speaker-net; 3xTF32 split (fp32-grade) encoder + "
                 "variance predictors" if args.policy else "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch, "phonemes": args.phonemes,
                   "mel_frames_per_step": frames_total, "L_max": L_max, "mel_frames_per_sec": frames_total / (ms / 1e3),
                   "audio_sec_per_step": audio_seconds(frames_total, cfg), "tensor_core_policy": args.policy,
                   "l2": "256 MiB buffer written between timed steps (L2 flush); activations per step also exceed L2",
                   "wall_s_timed_region": wall},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if e2e:
        line["e2e"] = {"value": audio_seconds(frames_total, cfg) / (e2e["ms"] / 1e3), "unit": UNIT,
                       "ms_per_step": e2e["ms"], "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"]}

    # ---- roofline of the dominant kernel class (CUDA events around every launch of the class, one extra step) ----
    peaks = measured_peaks()
    eng.profile(True)
    step_device()
    prof = eng.profile_read()
    eng.profile(False)
    dom = max(prof, key=lambda k: prof[k]["ms"])
    d = prof[dom]
    if d["launches"] > 0 and d["ms"] > 0:
        ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
        # TF32 tensor peak is half the dense bf16 rate; the measured denominators are bf16 (MEASURED_PEAKS.json)
        peak = peaks["bf16_tflops_sustained"]
        traffic, traffic_src = ncu_traffic(dom)
        line["roofline"] = {
            "bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "frac_of_tf32_peak": ach / (peak / 2),   # the kernel computes in TF32: half the bf16 rate the measured peak is for
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": d["bytes"] / d["launches"], "peak_source": f"bf16 dense, sustained, {peaks['source']} (MEASURED_PEAKS.json); "
                                            "TF32 runs at half the bf16 rate, fp32 FMA kernels at ~1/20",
            "launches_per_step": d["launches"], "avg_launch_ms": d["ms"] / d["launches"],
            "algorithmic_flops_per_launch": d["flops"] / d["launches"], "share_of_step": d["ms"] / ms,
            "pipeline": pipeline_roofline(prof, peaks, ms, (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0),
            "classes": {k: {"ms": v["ms"], "launches": v["launches"],
                            "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else None,
                            "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None} for k, v in prof.items()},
        }

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample, rank 0 at N = 1 only -----------
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        nb = min(args.cpu_sample_batch, args.batch)
        xs = {k: v[:nb].clone() for k, v in syn.make_inputs(cfg, args.batch, args.phonemes, args.ref_frames, seed=7).items()}
        t_cpu, f_cpu = cpu_reference_run(cfg, w, xs, threads, runs=2)
        line["cpu_baseline"] = {"value": audio_seconds(f_cpu, cfg) / t_cpu, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{nb} of the {args.batch} utterances of the same batch, median of 2 runs after "
                                          f"1 warm-up, torch.set_num_threads({threads})", "seconds_per_run": t_cpu}
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
