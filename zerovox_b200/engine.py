"""Tensor-level host API over the C ABI: torch supplies device memory and streams, nothing else.

Every method takes / returns CUDA tensors and enqueues work on torch's current stream.  There is no CPU
path: CPU tensors are rejected (copy them with ``.to(device)`` first, as the reference's callers do).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Mapping, Sequence

import torch

from . import _lib


@dataclass
class EngineConfig:
    """Hyper-parameters (ZeroVox.__init__ kwargs, model.py:159-201 + HiFi-GAN config.json, hifigan.py:93-110)."""
    num_phones: int = 28
    num_puncts: int = 10
    emb_dim: int = 512
    punct_emb_dim: int = 16
    max_txt_len: int = 512
    max_mel_len: int = 1750
    enc_layers: int = 4
    enc_heads: int = 2
    vp_filter_size: int = 256
    vp_kernel_size: int = 3
    ve_n_bins: int = 256
    decoder_kind: str = "fastspeech2"
    dec_layers: int = 6
    dec_heads: int = 2
    conv_filter_size: int = 1024
    conv_kernel_size: Sequence[int] = (9, 1)
    dec_scln: bool = True
    resnet_layers: Sequence[int] = (3, 4, 6, 3)
    resnet_num_filters: Sequence[int] = (32, 64, 128, 256)
    resnet_encoder_type: str = "ASP"
    n_mels: int = 80
    hop_length: int = 256
    hg_resblock: str = "1"
    hg_upsample_rates: Sequence[int] = (8, 8, 2, 2)
    hg_upsample_kernel_sizes: Sequence[int] = (16, 16, 4, 4)
    hg_upsample_initial_channel: int = 128
    hg_resblock_kernel_sizes: Sequence[int] = (3, 7, 11)
    hg_resblock_dilation_sizes: Sequence[Sequence[int]] = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    tensor_core_policy: int = 1

    @property
    def hidden(self) -> int:
        return self.emb_dim + self.punct_emb_dim

    def set_hifigan(self, h: Mapping) -> "EngineConfig":
        """Fill the vocoder fields from a HiFi-GAN config.json mapping / AttrDict."""
        self.hg_resblock = str(h["resblock"])
        self.hg_upsample_rates = tuple(h["upsample_rates"])
        self.hg_upsample_kernel_sizes = tuple(h["upsample_kernel_sizes"])
        self.hg_upsample_initial_channel = int(h["upsample_initial_channel"])
        self.hg_resblock_kernel_sizes = tuple(h["resblock_kernel_sizes"])
        self.hg_resblock_dilation_sizes = tuple(tuple(d) for d in h["resblock_dilation_sizes"])
        hop = 1
        for u in self.hg_upsample_rates:
            hop *= u
        self.hop_length = hop
        return self

    def to_c(self) -> _lib.ZvxConfig:
        c = _lib.ZvxConfig()
        c.abi_version = _lib.ZVX_ABI_VERSION
        for name in ("num_phones", "num_puncts", "emb_dim", "punct_emb_dim", "max_txt_len", "max_mel_len",
                     "enc_layers", "enc_heads", "vp_filter_size", "vp_kernel_size", "ve_n_bins", "dec_layers",
                     "dec_heads", "conv_filter_size", "n_mels", "hop_length", "hg_upsample_initial_channel",
                     "tensor_core_policy"):
            setattr(c, name, int(getattr(self, name)))
        kinds = {"fastspeech2": 0, "styletts": 1}
        if self.decoder_kind not in kinds:
            raise Exception(f"unknown decoder kind: '{self.decoder_kind}'")  # mirrors model.py:244
        c.decoder_kind = kinds[self.decoder_kind]
        c.dec_scln = 1 if self.dec_scln else 0
        c.conv_kernel_size[0], c.conv_kernel_size[1] = (int(k) for k in self.conv_kernel_size)
        if len(self.resnet_layers) != 4 or len(self.resnet_num_filters) != 4:
            raise ValueError("resnet_layers / resnet_num_filters need 4 entries")
        for i in range(4):
            c.resnet_layers[i] = int(self.resnet_layers[i])
            c.resnet_num_filters[i] = int(self.resnet_num_filters[i])
        enc = {"SAP": 0, "ASP": 1}
        if self.resnet_encoder_type not in enc:
            raise ValueError("Undefined encoder")  # mirrors ResNetSE34V2.py:151
        c.resnet_encoder_type = enc[self.resnet_encoder_type]
        c.hg_resblock = int(self.hg_resblock)
        nu, nk = len(self.hg_upsample_rates), len(self.hg_resblock_kernel_sizes)
        if nu > _lib.ZVX_MAX_UPSAMPLES or nk > _lib.ZVX_MAX_RESBLOCK_KERNELS:
            raise ValueError("HiFi-GAN config too large for the engine")
        c.hg_num_upsamples = nu
        for i in range(nu):
            c.hg_upsample_rates[i] = int(self.hg_upsample_rates[i])
            c.hg_upsample_kernel_sizes[i] = int(self.hg_upsample_kernel_sizes[i])
        c.hg_num_kernels = nk
        nd = len(self.hg_resblock_dilation_sizes[0])
        if nd > _lib.ZVX_MAX_DILATIONS or any(len(d) != nd for d in self.hg_resblock_dilation_sizes):
            raise ValueError("HiFi-GAN dilation table not supported")
        c.hg_num_dilations = nd
        for j in range(nk):
            c.hg_resblock_kernel_sizes[j] = int(self.hg_resblock_kernel_sizes[j])
            for d in range(nd):
                c.hg_resblock_dilation_sizes[j][d] = int(self.hg_resblock_dilation_sizes[j][d])
        return c


def _ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


class _nvtx:
    """NVTX range around one stage call: `nsys` / `ncu --nvtx` timelines show zvx_spkemb / zvx_encode /
    zvx_length_regulate / zvx_decode / zvx_vocode per forward.  Costs two driver calls; no-op when NVTX is unavailable."""
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        try:
            torch.cuda.nvtx.range_push(self.name)
        except Exception:  # noqa: BLE001
            self.name = None

    def __exit__(self, *exc):
        if self.name is not None:
            torch.cuda.nvtx.range_pop()


class Engine:
    """One engine handle bound to one CUDA device (not thread-safe, like the reference)."""

    def __init__(self, cfg: EngineConfig, device: torch.device | str | int = "cuda"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("zerovox_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if self.device.type != "cuda":
            raise RuntimeError(f"zerovox_b200 cannot run on device {self.device}; there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.cfg = cfg
        self._c_cfg = cfg.to_c()
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):   # the C side does cudaSetDevice: keep the caller's current device untouched
            rc = self.lib.zvx_create(C.byref(self._c_cfg), self.device.index, C.byref(self._h))
        if rc != 0:
            raise RuntimeError("zvx_create: " + self.lib.zvx_last_error(None).decode())

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self.lib.zvx_destroy(h)
            self._h = None

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what}: " + self.lib.zvx_last_error(self._h).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name}: expected a torch.Tensor")
        if t.device != self.device:
            raise RuntimeError(f"{name}: tensor is on {t.device}, engine is on {self.device} (no CPU path)")
        return t.to(dtype).contiguous()

    # ------------------------------------------------------------------ weights
    def set_weights(self, state_dict: Mapping[str, torch.Tensor], prefix: str = ""):
        """One zvx_set_weight per floating-point state_dict entry (key = prefix + name); host or device tensors."""
        for k, v in state_dict.items():
            if not isinstance(v, torch.Tensor) or not v.is_floating_point():
                continue
            t = v.detach().to(torch.float32).contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            with torch.cuda.device(self.device):
                self._check(self.lib.zvx_set_weight(self._h, (prefix + k).encode(), _ptr(t), shape, t.dim()),
                            f"zvx_set_weight({prefix + k})")

    def load_weights(self, state_dict: Mapping[str, torch.Tensor], prefix: str = ""):
        """set_weights + finalize (the engine-level load_state_dict)."""
        self.set_weights(state_dict, prefix)
        self.finalize()

    def finalize(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.zvx_finalize_weights(self._h), "zvx_finalize_weights")

    # ------------------------------------------------------------------ stages
    def spkemb(self, ref_mel: torch.Tensor) -> torch.Tensor:
        """ResNetSE34V2.forward: [B, T_ref, n_mels] -> [B, 1, hidden]."""
        x = self._dev(ref_mel, torch.float32, "ref_mel")
        B, T, M = x.shape
        if M != self.cfg.n_mels:
            raise RuntimeError(f"ref_mel has {M} mel channels, model expects {self.cfg.n_mels}")
        out = torch.empty((B, 1, self.cfg.hidden), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device), _nvtx("zvx_spkemb"):
            self._check(self.lib.zvx_spkemb(self._h, _ptr(x), B, T, _ptr(out), self._stream()), "zvx_spkemb")
        return out

    def spkemb_encode(self, ref_mel, phoneme, puncts, phoneme_mask=None, forced_duration=None):
        """``spkemb`` followed by ``encode`` as one C-ABI call (zvx_spkemb_encode): the speaker net runs on an engine-owned side
        stream next to the encoder's FFT blocks and joins the caller's stream where the style vector is first needed.  Returns
        (style [B, 1, hidden], the dict of :meth:`encode` with ``L_max`` / ``mel_len_host``)."""
        x = self._dev(ref_mel, torch.float32, "ref_mel")
        B, Tr, M = x.shape
        if M != self.cfg.n_mels:
            raise RuntimeError(f"ref_mel has {M} mel channels, model expects {self.cfg.n_mels}")
        ph = self._dev(phoneme, torch.int32, "phoneme")
        pu = self._dev(puncts, torch.int32, "puncts")
        if ph.shape[0] != B:
            raise RuntimeError("ref_mel batch size does not match phoneme batch size")
        T = ph.shape[1]
        pm = None if phoneme_mask is None else self._dev(phoneme_mask, torch.uint8, "phoneme_mask")
        fd = None if forced_duration is None else self._dev(forced_duration, torch.int32, "duration")
        dev, f32 = self.device, torch.float32
        style = torch.empty((B, 1, self.cfg.hidden), device=dev, dtype=f32)
        out = {
            "pitch": torch.empty((B, T), device=dev, dtype=f32),
            "energy": torch.empty((B, T), device=dev, dtype=f32),
            "log_duration": torch.empty((B, T), device=dev, dtype=f32),
            "duration_rounded": torch.empty((B, T), device=dev, dtype=torch.int32),
            "mel_len": torch.empty((B,), device=dev, dtype=torch.int64),
            "xprime": torch.empty((B, T, self.cfg.hidden), device=dev, dtype=f32),
        }
        lmax = C.c_int(0)
        host = (C.c_int64 * B)()
        with torch.cuda.device(self.device), _nvtx("zvx_spkemb_encode"):
            self._check(self.lib.zvx_spkemb_encode(
                self._h, _ptr(x), Tr, _ptr(style), _ptr(ph), _ptr(pu), _ptr(pm), _ptr(fd), B, T, _ptr(out["pitch"]),
                _ptr(out["energy"]), _ptr(out["log_duration"]), _ptr(out["duration_rounded"]), _ptr(out["mel_len"]),
                _ptr(out["xprime"]), C.cast(host, C.c_void_p), C.byref(lmax), self._stream()), "zvx_spkemb_encode")
        out["L_max"] = int(lmax.value)
        out["mel_len_host"] = list(host)
        return style, out

    def encode(self, phoneme, puncts, style, phoneme_mask=None, forced_duration=None, need_lengths=True):
        """FS2Encoder.forward up to the LengthRegulator.  Returns a dict of device tensors plus ``L_max`` (int)
        and ``mel_len_host`` (list[int]) when ``need_lengths`` (costs the engine's single stream sync)."""
        ph = self._dev(phoneme, torch.int32, "phoneme")
        pu = self._dev(puncts, torch.int32, "puncts")
        B, T = ph.shape
        st = self._dev(style, torch.float32, "style_embed").reshape(-1, self.cfg.hidden)
        if st.shape[0] == 1 and B > 1:
            st = st.expand(B, -1).contiguous()
        if st.shape[0] != B:
            raise RuntimeError("style_embed batch size does not match phoneme batch size")
        pm = None if phoneme_mask is None else self._dev(phoneme_mask, torch.uint8, "phoneme_mask")
        fd = None if forced_duration is None else self._dev(forced_duration, torch.int32, "duration")
        dev, f32 = self.device, torch.float32
        out = {
            "pitch": torch.empty((B, T), device=dev, dtype=f32),
            "energy": torch.empty((B, T), device=dev, dtype=f32),
            "log_duration": torch.empty((B, T), device=dev, dtype=f32),
            "duration_rounded": torch.empty((B, T), device=dev, dtype=torch.int32),
            "mel_len": torch.empty((B,), device=dev, dtype=torch.int64),
            "xprime": torch.empty((B, T, self.cfg.hidden), device=dev, dtype=f32),
        }
        lmax = C.c_int(0)
        host = (C.c_int64 * B)()
        with torch.cuda.device(self.device), _nvtx("zvx_encode"):
            self._check(self.lib.zvx_encode(
                self._h, _ptr(ph), _ptr(pu), _ptr(pm), _ptr(st), _ptr(fd), B, T, _ptr(out["pitch"]),
                _ptr(out["energy"]), _ptr(out["log_duration"]), _ptr(out["duration_rounded"]), _ptr(out["mel_len"]),
                _ptr(out["xprime"]), C.cast(host, C.c_void_p) if need_lengths else None,
                C.byref(lmax) if need_lengths else None, self._stream()), "zvx_encode")
        if need_lengths:
            out["L_max"] = int(lmax.value)
            out["mel_len_host"] = list(host)
        return out

    def length_regulate(self, xprime, duration, L_max: int, want_index=False, frame0: int = 0):
        """LengthRegulator.forward + pad: ([B,T,H], int32 [B,T]) -> features [B,L_max,H] (, src_index).  With ``frame0``
        only the frames [frame0, frame0 + L_max) are produced (chunked long-form processing)."""
        x = self._dev(xprime, torch.float32, "xprime")
        d = self._dev(duration, torch.int32, "duration")
        B, T, H = x.shape
        feats = torch.empty((B, L_max, H), device=self.device, dtype=torch.float32)
        idx = torch.empty((B, L_max), device=self.device, dtype=torch.int32) if want_index else None
        with torch.cuda.device(self.device), _nvtx("zvx_length_regulate"):
            if frame0:
                self._check(self.lib.zvx_length_regulate_chunk(self._h, _ptr(x), _ptr(d), B, T, int(frame0), L_max,
                                                               _ptr(feats), _ptr(idx), self._stream()),
                            "zvx_length_regulate_chunk")
            else:
                self._check(self.lib.zvx_length_regulate(self._h, _ptr(x), _ptr(d), B, T, L_max, _ptr(feats), _ptr(idx),
                                                         self._stream()), "zvx_length_regulate")
        return (feats, idx) if want_index else feats

    def vocode_chunked(self, mel_bcl: torch.Tensor, chunk_frames: int = 512, halo_frames: int = 14) -> torch.Tensor:
        """hifigan.Generator.forward over a long mel in chunks of ``chunk_frames`` with ``halo_frames`` of context on
        each side, keeping only the samples of the chunk proper (overlap-discard): exact, because the generator's
        receptive field is < 14 mel frames per side for every upstream HiFi-GAN topology, while
        the workspace stays bounded by the chunk size.  [B, n_mels, L] -> [B, 1, L*hop]."""
        m = self._dev(mel_bcl, torch.float32, "mel")
        B, Cm, L = m.shape
        hop = self.cfg.hop_length
        wav = torch.empty((B, 1, L * hop), device=self.device, dtype=torch.float32)
        for c0 in range(0, L, chunk_frames):
            c1 = min(L, c0 + chunk_frames)
            a, b = max(0, c0 - halo_frames), min(L, c1 + halo_frames)
            part = self.vocode(m[:, :, a:b].contiguous())
            wav[:, :, c0 * hop:c1 * hop] = part[:, :, (c0 - a) * hop:(c1 - a) * hop]
        return wav

    def decode(self, features, style, mask=None, mel_len=None, zero_padded_mel=False, want_blc=True, want_bcl=True):
        """FS2Decoder.forward (+ model.py:283-285 masking).  Returns (mel [B,L,n_mels] | None, mel [B,n_mels,L] | None)."""
        f = self._dev(features, torch.float32, "features")
        B, L, H = f.shape
        st = self._dev(style, torch.float32, "spk_emb").reshape(-1, self.cfg.hidden)
        if st.shape[0] == 1 and B > 1:
            st = st.expand(B, -1).contiguous()
        m = None if mask is None else self._dev(mask, torch.uint8, "mask")
        ml = None if mel_len is None else self._dev(mel_len, torch.int64, "mel_len")
        blc = torch.empty((B, L, self.cfg.n_mels), device=self.device, dtype=torch.float32) if want_blc else None
        bcl = torch.empty((B, self.cfg.n_mels, L), device=self.device, dtype=torch.float32) if want_bcl else None
        with torch.cuda.device(self.device), _nvtx("zvx_decode"):
            self._check(self.lib.zvx_decode(self._h, _ptr(f), _ptr(m), _ptr(ml), _ptr(st), B, L,
                                            1 if zero_padded_mel else 0, _ptr(blc), _ptr(bcl), self._stream()),
                        "zvx_decode")
        return blc, bcl

    def vocode(self, mel_bcl: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """hifigan.Generator.forward: [B, n_mels, L] -> [B, 1, L*hop] (written into ``out`` when given: a contiguous
        [B, 1, L*hop] or [B, L*hop] fp32 tensor, e.g. a batch slice of a larger result)."""
        m = self._dev(mel_bcl, torch.float32, "mel")
        B, Cm, L = m.shape
        if Cm != self.cfg.n_mels:
            raise RuntimeError(f"mel has {Cm} channels, vocoder expects {self.cfg.n_mels}")
        if out is None:
            wav = torch.empty((B, 1, L * self.cfg.hop_length), device=self.device, dtype=torch.float32)
        else:
            if out.device != self.device or out.dtype != torch.float32 or not out.is_contiguous() or \
                    out.numel() != B * L * self.cfg.hop_length:
                raise RuntimeError("vocode(out=): need a contiguous fp32 tensor of B * L * hop elements on the engine's device")
            wav = out
        with torch.cuda.device(self.device), _nvtx("zvx_vocode"):
            self._check(self.lib.zvx_vocode(self._h, _ptr(m), B, L, _ptr(wav), self._stream()), "zvx_vocode")
        return wav

    # ------------------------------------------------------------------ introspection
    PROF_CLASSES = {"gemm_fp32": 0, "gemm_tf32_tcgen05": 1, "vocoder_conv1d": 2, "vocoder_upsample": 3,
                    "vocoder_pair_tcgen05": 4, "gemm_3xtf32_tcgen05": 5}

    def profile(self, on: bool):
        self._check(self.lib.zvx_profile_enable(self._h, 1 if on else 0), "zvx_profile_enable")

    def profile_read(self) -> dict:
        """{class: {ms, launches, flops, bytes}} for the launches recorded since profile(True) (synchronises)."""
        out = {}
        for name, cls in self.PROF_CLASSES.items():
            ms, n, fl, by = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
            self._check(self.lib.zvx_profile_read(self._h, cls, C.byref(ms), C.byref(n), C.byref(fl), C.byref(by)),
                        "zvx_profile_read")
            out[name] = {"ms": ms.value, "launches": n.value, "flops": fl.value, "bytes": by.value}
        return out

    def set_option(self, name: str, value: int):
        """zvx_set_option: e.g. ``score_workspace_bytes`` (attention-score budget; long inputs are chunked over query rows)."""
        self._check(self.lib.zvx_set_option(self._h, name.encode(), int(value)), "zvx_set_option")

    def workspace_bytes(self) -> int:
        return int(self.lib.zvx_workspace_bytes(self._h))

    def launch_count(self) -> int:
        return int(self.lib.zvx_launch_count(self._h))
