"""HiFi-GAN Generator with the reference's interface (zerovox/tts/hifigan.py:89-139): constructed from the
config.json AttrDict, loads ``generator.ckpt['generator']`` (weight_g / weight_v form), ``remove_weight_norm()``,
``forward(mel [B,80,L] | [80,L]) -> [B,1,256L] | [1,256L]``.  The arithmetic runs in zvx_vocode."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn.utils import weight_norm, remove_weight_norm

from ._context import EngineModuleMixin


def _same_pad(k, d=1):
    return (k * d - d) // 2


class _ResBlockParams(nn.Module):
    def __init__(self, kind, ch, k, dilations):
        super().__init__()
        mk = lambda d: weight_norm(nn.Conv1d(ch, ch, k, 1, dilation=d, padding=_same_pad(k, d)))
        if kind == "1":   # keys convs1.{i}, convs2.{i}   (hifigan.py:25-47)
            self.convs1 = nn.ModuleList(mk(d) for d in dilations)
            self.convs2 = nn.ModuleList(mk(1) for _ in dilations)
        else:             # keys convs.{i}   (hifigan.py:65-76)
            self.convs = nn.ModuleList(mk(d) for d in dilations)

    def remove_weight_norm(self):
        for m in self.modules():
            if isinstance(m, nn.Conv1d) and hasattr(m, "weight_g"):
                remove_weight_norm(m)


class Generator(EngineModuleMixin, nn.Module):
    _role = "vocoder"

    def __init__(self, h):
        super().__init__()
        self.h = h
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        c0 = h.upsample_initial_channel
        self.conv_pre = weight_norm(nn.Conv1d(80, c0, 7, 1, padding=3))
        self.ups = nn.ModuleList(
            weight_norm(nn.ConvTranspose1d(c0 // (2 ** i), c0 // (2 ** (i + 1)), k, u, padding=(k - u) // 2))
            for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)))
        self.resblocks = nn.ModuleList()
        ch = c0
        for i in range(self.num_upsamples):
            ch = c0 // (2 ** (i + 1))
            for k, d in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                self.resblocks.append(_ResBlockParams(str(h.resblock), ch, k, d))
        self.conv_post = weight_norm(nn.Conv1d(ch, 1, 7, 1, padding=3))
        self._init_engine_binding()

    def remove_weight_norm(self):
        for m in list(self.ups) + [self.conv_pre, self.conv_post]:
            if hasattr(m, "weight_g"):
                remove_weight_norm(m)
        for r in self.resblocks:
            r.remove_weight_norm()
        self._ctx.mark_stale()

    def _fill_config(self, cfg):
        cfg.set_hifigan(self.h)

    def _engine_state_dict(self):
        """Plain ``weight`` for every conv: folds g * v / ||v|| (dim 0) when weight norm is still attached."""
        sd = {}
        for name, m in self.named_modules():
            if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d)):
                w = torch._weight_norm(m.weight_v, m.weight_g, 0) if hasattr(m, "weight_g") else m.weight
                sd[name + ".weight"] = w.detach()
                sd[name + ".bias"] = m.bias.detach()
        return sd

    def forward(self, x):
        eng = self._engine()
        unbatched = x.dim() == 2
        wav = eng.vocode((x.unsqueeze(0) if unbatched else x).to(eng.device))
        return wav.squeeze(0) if unbatched else wav

    def forward_chunked(self, x, chunk_frames=512, halo_frames=14):
        """Long-form vocoding (BASELINE config 5): same result as ``forward`` computed chunk by chunk with a
        ``halo_frames`` context on each side that is discarded afterwards (exact overlap-discard, bounded workspace)."""
        eng = self._engine()
        unbatched = x.dim() == 2
        wav = eng.vocode_chunked((x.unsqueeze(0) if unbatched else x).to(eng.device), chunk_frames, halo_frames)
        return wav.squeeze(0) if unbatched else wav
