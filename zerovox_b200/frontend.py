"""Host API of the callers either side of the path: the speaker-prompt front-end (silence
trim + log-mel spectrogram on the GPU) and the tokeniser / padding collator (host C code), over the C ABI.

No CPU path for the audio work: :class:`MelFrontend` raises without a CUDA device.  The tokeniser is host code by
nature (as in the reference) but still lives in the native library — there is no Python re-implementation behind it.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device):
    """torch's current stream ON THE HANDLE'S DEVICE (not on whatever device happens to be current)."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class MelFrontend:
    """trim + get_mel_from_wav of ZeroVoxTTS.speaker_embed (synthesize.py:123-138; mels.py:356-394) on one GPU."""

    def __init__(self, sampling_rate=22050, fft_size=1024, hop_size=256, win_length=1024, num_mels=80, fmin=0,
                 fmax=8000, device: torch.device | str | int = "cuda"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("zerovox_b200 mel front-end needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if self.device.type != "cuda":
            raise RuntimeError(f"zerovox_b200 cannot run on device {self.device}; there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        c = _lib.ZvxMelConfig()
        c.abi_version = _lib.ZVX_ABI_VERSION
        self.sampling_rate = int(sampling_rate)
        c.sampling_rate, c.fft_size, c.hop_size = int(sampling_rate), int(fft_size), int(hop_size)
        c.win_length = int(win_length if win_length is not None else fft_size)
        c.num_mels, c.fmin, c.fmax, c.clip_val = int(num_mels), float(fmin or 0), float(fmax), 1e-5
        self.num_mels, self.hop_size = c.num_mels, c.hop_size
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):   # the C side does cudaSetDevice: keep the caller's current device untouched
            rc = self.lib.zvx_frontend_create(C.byref(c), self.device.index, C.byref(self._h))
        if rc != 0:
            raise RuntimeError("zvx_frontend_create: " + self.lib.zvx_frontend_last_error(None).decode())

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self.lib.zvx_frontend_destroy(h)
            self._h = None

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: " + self.lib.zvx_frontend_last_error(self._h).decode())

    def _wav(self, wav: torch.Tensor) -> torch.Tensor:
        if not isinstance(wav, torch.Tensor) or wav.device != self.device:
            raise RuntimeError(f"expected a tensor on {self.device} (there is no CPU path)")
        if wav.dim() == 1:
            wav = wav[None]
        return wav.contiguous().float()

    def num_frames(self, n_samples: int) -> int:
        return int(self.lib.zvx_mel_num_frames(self._h, int(n_samples)))

    def trim(self, wav: torch.Tensor, wav_len: torch.Tensor | None = None, top_db: float = 40.0,
             frame_length: int = 2048, hop_length: int = 512):
        """librosa.effects.trim per row of wav [B, n] -> (start, length) device int64 [B] + the same as host lists
        (one stream sync, to size what follows)."""
        wav = self._wav(wav)
        B, n = wav.shape
        start = torch.empty(B, dtype=torch.int64, device=self.device)
        length = torch.empty(B, dtype=torch.int64, device=self.device)
        host = (C.c_int64 * (2 * B))()
        with torch.cuda.device(self.device):
            self._check(self.lib.zvx_trim_silence(self._h, _ptr(wav), B, n, _ptr(wav_len), float(top_db), int(frame_length),
                                                  int(hop_length), _ptr(start), _ptr(length), host, _stream(self.device)),
                        "zvx_trim_silence")
        return start, length, list(host[:B]), list(host[B:])

    def mel(self, wav: torch.Tensor, wav_start: torch.Tensor | None = None, wav_len: torch.Tensor | None = None,
            n_frames: int | None = None, with_energy: bool = False):
        """get_mel_from_wav per row -> mel [B, n_frames, num_mels] (the `_spkemb` layout), optional energy [B, n_frames].
        Without `n_frames` the whole row length decides (no window arguments) — pass it when rows are windowed."""
        wav = self._wav(wav)
        B, n = wav.shape
        if n_frames is None:
            if wav_start is not None or wav_len is not None:
                raise ValueError("pass n_frames together with wav_start / wav_len")
            n_frames = self.num_frames(n)
        mel = torch.empty((B, n_frames, self.num_mels), dtype=torch.float32, device=self.device)
        energy = torch.empty((B, n_frames), dtype=torch.float32, device=self.device) if with_energy else None
        with torch.cuda.device(self.device):
            self._check(self.lib.zvx_mel_spectrogram(self._h, _ptr(wav), B, n, _ptr(wav_start), _ptr(wav_len), int(n_frames),
                                                     _ptr(mel), _ptr(energy), _stream(self.device)), "zvx_mel_spectrogram")
        return (mel, energy) if with_energy else mel

    def resample(self, wav: torch.Tensor, sr_in: int, sr_out: int | None = None, wav_len: torch.Tensor | None = None):
        """The `sr=` conversion of librosa.load (synthesize.py:113-121) on the GPU: wav [B, n] (or [n]) at ``sr_in`` ->
        [B, ceil(n * sr_out / sr_in)] at ``sr_out`` (default: this front-end's sampling rate)."""
        sr_out = int(sr_out or self.sampling_rate)
        wav = self._wav(wav)
        if int(sr_in) == sr_out:
            return wav
        B, n = wav.shape
        n_out = int(self.lib.zvx_resample_num_samples(n, int(sr_in), sr_out))
        out = torch.empty((B, n_out), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.zvx_resample(self._h, _ptr(wav), B, n, _ptr(wav_len), int(sr_in), sr_out, _ptr(out), n_out, n_out,
                                              _stream(self.device)), "zvx_resample")
        return out

    def speaker_prompt_mel(self, wav: torch.Tensor, top_db: float = 40.0) -> torch.Tensor:
        """synthesize.py:123-138 for a batch of prompts [B, n]: trim each row, then its log-mel; rows are zero-filled
        beyond their own frame count.  Returns [B, max frames, num_mels] ready for `_spkemb`."""
        start, length, _, hlen = self.trim(wav, top_db=top_db)
        n_frames = max(self.num_frames(n) for n in hlen)
        return self.mel(wav, wav_start=start, wav_len=length, n_frames=n_frames)


class Tokeniser:
    """Symbols + ZeroVoxTTS.transcript2phonemids + collate_fn padding (symbols.py:2-49; synthesize.py:145-190;
    data.py:56-83) in the native library."""

    def __init__(self, phones: str, puncts: str):
        self.lib = _lib.load()
        self._h = C.c_void_p()
        rc = self.lib.zvx_symbols_create(phones.encode("utf-8"), puncts.encode("utf-8"), C.byref(self._h))
        if rc != 0:
            raise RuntimeError("zvx_symbols_create: " + self.lib.zvx_symbols_last_error(None).decode())

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self.lib.zvx_symbols_destroy(h)
            self._h = None

    @property
    def num_phones(self) -> int:
        return self.lib.zvx_symbols_num_phones(self._h)

    @property
    def num_puncts(self) -> int:
        return self.lib.zvx_symbols_num_puncts(self._h)

    def transcript2phonemids(self, transcript: str) -> tuple[list[int], list[int]]:
        if "\x00" in transcript:      # C strings end at NUL; the reference would skip it as a non-phone
            transcript = transcript.replace("\x00", "")
        raw = transcript.encode("utf-8")
        cap = max(1, len(transcript))
        ph = (C.c_int32 * cap)()
        pu = (C.c_int32 * cap)()
        n = self.lib.zvx_transcript2phonemids(self._h, raw, ph, pu, cap)
        if n == -2:
            raise KeyError(" ")       # the reference's encode_punct(' ') (symbols.py:41-42)
        if n < 0 or n > cap:
            raise RuntimeError("zvx_transcript2phonemids: " + self.lib.zvx_symbols_last_error(self._h).decode())
        return list(ph[:n]), list(pu[:n])

    def collate(self, phone_seqs: Sequence[Sequence[int]], punct_seqs: Sequence[Sequence[int]], pinned: bool = False):
        """-> phoneme i32 [B,T], puncts i32 [B,T], phoneme_mask bool [B,T] (True = padding), lens i32 [B] (CPU tensors,
        optionally pinned, ready for one H2D copy each)."""
        B = len(phone_seqs)
        lens = np.array([len(s) for s in phone_seqs], dtype=np.int32)
        if any(len(q) != l for q, l in zip(punct_seqs, lens)):
            raise ValueError("phone and punct sequences differ in length")
        T = int(lens.max()) if B else 0
        arrs_ph = [np.ascontiguousarray(s, dtype=np.int32) for s in phone_seqs]
        arrs_pu = [np.ascontiguousarray(s, dtype=np.int32) for s in punct_seqs]
        pp = (C.c_void_p * max(B, 1))(*[a.ctypes.data for a in arrs_ph])
        qq = (C.c_void_p * max(B, 1))(*[a.ctypes.data for a in arrs_pu])
        pin = pinned and torch.cuda.is_available()
        phoneme = torch.empty((B, T), dtype=torch.int32, pin_memory=pin)
        puncts = torch.empty((B, T), dtype=torch.int32, pin_memory=pin)
        mask = torch.empty((B, T), dtype=torch.uint8, pin_memory=pin)
        rc = self.lib.zvx_collate(pp, qq, C.c_void_p(lens.ctypes.data), B, T, _ptr(phoneme), _ptr(puncts), _ptr(mask))
        if rc != 0:
            raise RuntimeError(f"zvx_collate failed ({rc})")
        return phoneme, puncts, mask.view(torch.bool), torch.from_numpy(lens)
