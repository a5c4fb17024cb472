"""ResNetSE34V2 speaker-embedding net with the reference's interface and state_dict keys
(zerovox/tts/ResNetSE34V2.py:101-212): forward(ref_mel [B,T,n_mels]) -> [B,1,nOut] (unit L2 norm).
The arithmetic runs in zvx_spkemb."""
from __future__ import annotations

import torch
import torch.nn as nn

from ._context import EngineModuleMixin


class _SEParams(nn.Module):            # keys fc.0, fc.2
    def __init__(self, ch, reduction=8):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(ch, ch // reduction), nn.ReLU(), nn.Linear(ch // reduction, ch), nn.Sigmoid())


class _BlockParams(nn.Module):         # keys conv1, bn1, conv2, bn2, se.fc.{0,2}, downsample.{0,1}
    def __init__(self, inpl, planes, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(inpl, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.se = _SEParams(planes)
        if stride != 1 or inpl != planes:
            self.downsample = nn.Sequential(nn.Conv2d(inpl, planes, 1, stride=stride, bias=False),
                                            nn.BatchNorm2d(planes))


class _UnusedFrontEnd(nn.Module):
    """Buffers the reference registers but never uses in forward (PreEmphasis + MelSpectrogram,
    ResNetSE34V2.py:34-50, 123-126); kept so checkpoints load with identical keys."""

    def __init__(self, n_mels):
        super().__init__()
        pre = nn.Module()
        pre.register_buffer("flipped_filter", torch.tensor([[[-0.97, 1.0]]]))
        mel = nn.Module()
        mel.spectrogram = nn.Module()
        mel.spectrogram.register_buffer("window", torch.hamming_window(400))
        mel.mel_scale = nn.Module()
        mel.mel_scale.register_buffer("fb", torch.zeros(257, n_mels))
        self.add_module("0", pre)
        self.add_module("1", mel)


class ResNetSE34V2(EngineModuleMixin, nn.Module):
    _role = "spkemb"

    def __init__(self, layers, num_filters, nOut, encoder_type, n_mels, log_input):
        super().__init__()
        if log_input:
            raise NotImplementedError("zerovox_b200: log_input=True is never used by ZeroVox (model.py:223)")
        if encoder_type not in ("ASP", "SAP"):
            raise ValueError("Undefined encoder")
        self._hp = dict(resnet_layers=tuple(layers), resnet_num_filters=tuple(num_filters),
                        resnet_encoder_type=encoder_type, n_mels=n_mels)
        self._n_out = nOut
        self.encoder_type, self.n_mels, self.log_input = encoder_type, n_mels, log_input
        self.conv1 = nn.Conv2d(1, num_filters[0], 3, stride=1, padding=1)
        self.bn1 = nn.BatchNorm2d(num_filters[0])
        inpl = num_filters[0]
        for li, (planes, n) in enumerate(zip(num_filters, layers), start=1):
            blocks = []
            for bi in range(n):
                blocks.append(_BlockParams(inpl, planes, 2 if (li > 1 and bi == 0) else 1))
                inpl = planes
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        self.torchfb = _UnusedFrontEnd(n_mels)
        d = num_filters[3] * (n_mels // 8)
        self.attention = nn.Sequential(nn.Conv1d(d, 128, 1), nn.ReLU(), nn.BatchNorm1d(128), nn.Conv1d(128, d, 1),
                                       nn.Softmax(dim=2))
        self.fc = nn.Linear(d * 2 if encoder_type == "ASP" else d, nOut)
        self._init_engine_binding()

    def _fill_config(self, cfg):
        for k, v in self._hp.items():
            setattr(cfg, k, v)
        if cfg.hidden != self._n_out:
            cfg.emb_dim, cfg.punct_emb_dim = self._n_out - 16, 16

    def forward(self, x, l2_norm=True):
        if not l2_norm:
            raise NotImplementedError("zerovox_b200: l2_norm=False is never used by ZeroVox")
        eng = self._engine()
        return eng.spkemb(x.to(eng.device))
