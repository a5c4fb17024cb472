#!/usr/bin/env python
"""The reference's demo flow (zerovox/demo.py:60-135) on the B200 engine: load a model, embed a speaker prompt, speak a
sentence, write a .wav and report the real-time factor.

    python tools/demo.py --model /path/to/tts_en_model --meldec-model zerovox-hifigan-vctk-v2-en-1 \\
        --refaudio prompt.wav --wav-filename out.wav "this is a test."          # real checkpoints (local, nothing is downloaded)
    python tools/demo.py --synthetic --wav-filename out.wav --iter 20 "this is a test."   # seeded random weights + prompt

Differences from the reference CLI: `--infer-device` is always a CUDA device (no CPU path); the text must already be
normalised (lower-case phones of the model's alphabet — the NeMo / uroman normaliser is outside this engine, see
DESIGN.md §9); `--refaudio` is read with scipy (16-bit / float PCM at the model's sampling rate, no resampling); no audio
playback.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy.io import wavfile  # noqa: E402


class PassThroughNormalizer:
    """Stands in for ZeroVoxNormalizer (zerovox/tts/normalize.py): lower-cases, nothing else."""
    language = "en"

    def normalize(self, text):
        return text.lower(), None


def write_wav_to_file(wav, length, filename, sample_rate, hop_length):   # demo.py:27-33
    wav = (wav * 32760).astype("int16")
    wav = wav[: length * hop_length]
    print("Writing wav to {}".format(filename))
    wavfile.write(filename, sample_rate, wav)


def read_prompt(path, sampling_rate):
    sr, data = wavfile.read(path)
    if sr != sampling_rate:
        raise SystemExit(f"{path}: {sr} Hz, the model wants {sampling_rate} Hz (resampling is outside this engine)")
    if data.ndim > 1:
        data = data.mean(axis=1)
    if data.dtype.kind == "i":
        data = data.astype(np.float32) / float(np.iinfo(data.dtype).max)
    return data.astype(np.float32)


def main():
    p = argparse.ArgumentParser(prog="demo", description="zerovox demo flow on the B200 engine")
    p.add_argument("--infer-device", default="cuda:0")
    p.add_argument("--model", help="model directory (modelcfg.yaml + checkpoints/*.ckpt) or cached model name")
    p.add_argument("--meldec-model", default="zerovox-hifigan-vctk-v2-en-1")
    p.add_argument("--synthetic", action="store_true", help="seeded random weights (tts_medium_styledec + HiFi-GAN V2) and prompt")
    p.add_argument("--refaudio", help="reference audio .wav at the model's sampling rate")
    p.add_argument("--wav-filename", help=".wav file to produce")
    p.add_argument("--iter", type=int, default=1, help="iterations (for benchmarking), default: 1")
    p.add_argument("--verbose", action="store_true")
    p.add_argument("text", nargs="?", default="this is a test.")
    args = p.parse_args()

    from zerovox_b200.tts.synthesize import ZeroVoxTTS
    if args.synthetic or not args.model:
        import dataclasses
        from zerovox_b200 import synthetic as syn
        from zerovox_b200.testing import build_model
        from zerovox_b200.tts.symbols import Symbols
        cfg = dataclasses.replace(syn.ZeroVoxConfig(), decoder_kind="styletts")       # the shipped default models' decoder
        w = syn.make_weights(cfg, seed=0, dur_bias=float(np.log(7.0)))               # ~6 frames per phoneme
        model = build_model(cfg, w, device=args.infer_device)
        synth = ZeroVoxTTS(language="en", syms=Symbols(cfg.phones, cfg.puncts), checkpoint=None, meldec_model=None,
                           hop_length=cfg.hop_length, sampling_rate=cfg.sampling_rate, n_mel_channels=cfg.n_mels,
                           fft_size=1024, win_length=1024, mel_fmin=0, mel_fmax=8000, infer_device=args.infer_device,
                           verbose=args.verbose, model=model, normalizer=PassThroughNormalizer())
        sampling_rate, hop = cfg.sampling_rate, cfg.hop_length
        prompt = syn.make_speech_like(5 * sampling_rate, seed=1) if not args.refaudio else read_prompt(args.refaudio, sampling_rate)
    else:
        modelcfg, synth = ZeroVoxTTS.load_model(args.model, meldec_model=args.meldec_model, infer_device=args.infer_device,
                                                verbose=args.verbose, normalizer=PassThroughNormalizer())
        sampling_rate, hop = modelcfg["audio"]["sampling_rate"], modelcfg["audio"]["hop_size"]
        if not args.refaudio:
            raise SystemExit("--refaudio is required with --model")
        prompt = read_prompt(args.refaudio, sampling_rate)

    t0 = time.time()
    spkemb = synth.speaker_embed(prompt)
    torch.cuda.synchronize()
    print(f"speaker embedding of a {len(prompt) / sampling_rate:.2f} s prompt: {time.time() - t0:.3f} s (first call, includes weight upload)")

    rtf, warmup = [], min(10, max(0, args.iter - 1))
    for i in range(args.iter):
        torch.cuda.synchronize()
        start_time = time.time()
        wav, phoneme, length = synth.tts(args.text, spkemb)            # wav comes back as numpy: the call is synchronous
        elapsed_time = time.time() - start_time
        wav_len = wav.shape[0] / sampling_rate
        real_time_factor = wav_len / elapsed_time
        print(f"[{i + 1}/{args.iter}] Synth time: {elapsed_time * 1e3:.2f} ms, voice length: {wav_len:.2f} sec, rtf: {real_time_factor:.1f}")
        if i >= warmup:
            rtf.append(real_time_factor)
    if args.wav_filename:
        write_wav_to_file(wav, length=length, filename=args.wav_filename, sample_rate=sampling_rate, hop_length=hop)
    if rtf:
        print("Average RTF: {:.1f}".format(float(np.mean(rtf))))


if __name__ == "__main__":
    main()
