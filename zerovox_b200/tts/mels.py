"""Mirror of the inference-time part of zerovox/tts/mels.py: ``get_mel_from_wav`` (mels.py:356-394), computed by the
fused CUDA front-end (csrc/frontend.cu) instead of librosa.  Same argument names, same returns:
``(spec [num_mels, n_frames] float32, energy [n_frames] float32)`` as numpy arrays for numpy input.  There is no CPU
path: the call needs a CUDA device."""
from __future__ import annotations

import numpy as np
import torch

from ..frontend import MelFrontend

_frontends: dict = {}


def _frontend(device, sampling_rate, fft_size, hop_size, win_length, num_mels, fmin, fmax) -> MelFrontend:
    # the reference caches one global mel_basis and asserts fmax never changes (mels.py:376-382); here one front-end
    # per distinct parameter set and device
    key = (str(device), sampling_rate, fft_size, hop_size, win_length, num_mels, fmin, fmax)
    if key not in _frontends:
        _frontends[key] = MelFrontend(sampling_rate, fft_size, hop_size, win_length, num_mels, fmin, fmax, device=device)
    return _frontends[key]


def get_mel_from_wav(audio, sampling_rate, fft_size, hop_size, win_length, num_mels, fmin, fmax, device="cuda"):
    as_numpy = not isinstance(audio, torch.Tensor)
    if as_numpy:
        if np.min(audio) < -1.:
            print(f"WARNING: get_mel_from_wav: audio min value < -1.0 : {np.min(audio)}")
        if np.max(audio) > 1.:
            print(f"WARNING: get_mel_from_wav: audio max value >  1.0 : {np.max(audio)}")
        audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).to(device)
    fe = _frontend(audio.device, sampling_rate, fft_size, hop_size, win_length, num_mels, fmin, fmax)
    mel, energy = fe.mel(audio, with_energy=True)
    spec = mel[0].transpose(0, 1)          # [num_mels, n_frames], the reference's orientation
    if as_numpy:
        return spec.contiguous().cpu().numpy(), energy[0].cpu().numpy()
    return spec, energy[0]
