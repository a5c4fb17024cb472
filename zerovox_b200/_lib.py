"""ctypes binding of libzerovox_b200.so (the C ABI declared in include/zerovox_b200.h).

There is no fallback: if the CUDA library has not been built (``python -c "import __graft_entry__ as g; g.build()"``)
importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzerovox_b200.so")

ZVX_MAX_UPSAMPLES = 8
ZVX_MAX_RESBLOCK_KERNELS = 8
ZVX_MAX_DILATIONS = 4
ZVX_ABI_VERSION = 1


class ZvxConfig(C.Structure):
    """Mirror of ``struct zvx_config`` (include/zerovox_b200.h) — keep field order identical."""
    _fields_ = [
        ("abi_version", C.c_int32),
        ("num_phones", C.c_int32),
        ("num_puncts", C.c_int32),
        ("emb_dim", C.c_int32),
        ("punct_emb_dim", C.c_int32),
        ("max_txt_len", C.c_int32),
        ("max_mel_len", C.c_int32),
        ("enc_layers", C.c_int32),
        ("enc_heads", C.c_int32),
        ("vp_filter_size", C.c_int32),
        ("vp_kernel_size", C.c_int32),
        ("ve_n_bins", C.c_int32),
        ("decoder_kind", C.c_int32),
        ("dec_layers", C.c_int32),
        ("dec_heads", C.c_int32),
        ("conv_filter_size", C.c_int32),
        ("conv_kernel_size", C.c_int32 * 2),
        ("dec_scln", C.c_int32),
        ("resnet_layers", C.c_int32 * 4),
        ("resnet_num_filters", C.c_int32 * 4),
        ("resnet_encoder_type", C.c_int32),
        ("n_mels", C.c_int32),
        ("hop_length", C.c_int32),
        ("hg_resblock", C.c_int32),
        ("hg_num_upsamples", C.c_int32),
        ("hg_upsample_rates", C.c_int32 * ZVX_MAX_UPSAMPLES),
        ("hg_upsample_kernel_sizes", C.c_int32 * ZVX_MAX_UPSAMPLES),
        ("hg_upsample_initial_channel", C.c_int32),
        ("hg_num_kernels", C.c_int32),
        ("hg_resblock_kernel_sizes", C.c_int32 * ZVX_MAX_RESBLOCK_KERNELS),
        ("hg_num_dilations", C.c_int32),
        ("hg_resblock_dilation_sizes", (C.c_int32 * ZVX_MAX_DILATIONS) * ZVX_MAX_RESBLOCK_KERNELS),
        ("tensor_core_policy", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class ZvxMelConfig(C.Structure):
    """Mirror of ``struct zvx_mel_config`` (speaker-prompt front-end)."""
    _fields_ = [(n, C.c_int32) for n in ("abi_version", "sampling_rate", "fft_size", "hop_size", "win_length",
                                         "num_mels")] + \
               [(n, C.c_float) for n in ("fmin", "fmax", "clip_val")] + [("reserved", C.c_int32 * 7)]


class ZvxGemmDesc(C.Structure):
    """Mirror of ``struct zvx_gemm_desc`` (kernel-level test hook)."""
    _fields_ = [(n, C.c_void_p) for n in ("A", "W", "C", "bias", "scale", "shift", "R")] + \
               [(n, C.c_int32) for n in ("M", "N", "K", "taps", "mode", "L", "Hh", "Ww", "ksize", "pad", "dil",
                                         "relu_first", "relu_last", "lda", "ldw", "ldc", "stride")]


# every symbol include/zerovox_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "zvx_abi_version": (C.c_int, []),
    "zvx_create": (C.c_int, [C.POINTER(ZvxConfig), C.c_int, C.POINTER(_P)]),
    "zvx_destroy": (None, [_P]),
    "zvx_last_error": (C.c_char_p, [_P]),
    "zvx_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "zvx_finalize_weights": (C.c_int, [_P]),
    "zvx_spkemb": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "zvx_spkemb_encode": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P,
                                    C.POINTER(C.c_int), _P]),
    "zvx_encode": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P,
                             C.POINTER(C.c_int), _P]),
    "zvx_length_regulate": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "zvx_length_regulate_chunk": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "zvx_decode": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "zvx_vocode": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "zvx_debug_gemm": (C.c_int, [_P, C.POINTER(ZvxGemmDesc), C.c_int, _P]),
    "zvx_profile_enable": (C.c_int, [_P, C.c_int]),
    "zvx_profile_read": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                   C.POINTER(C.c_double)]),
    "zvx_frontend_create": (C.c_int, [C.POINTER(ZvxMelConfig), C.c_int, C.POINTER(_P)]),
    "zvx_frontend_destroy": (None, [_P]),
    "zvx_frontend_last_error": (C.c_char_p, [_P]),
    "zvx_mel_num_frames": (C.c_int64, [_P, C.c_int64]),
    "zvx_trim_silence": (C.c_int, [_P, _P, C.c_int, C.c_int64, _P, C.c_float, C.c_int, C.c_int, _P, _P, _P, _P]),
    "zvx_mel_spectrogram": (C.c_int, [_P, _P, C.c_int, C.c_int64, _P, _P, C.c_int, _P, _P, _P]),
    "zvx_resample_num_samples": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "zvx_resample": (C.c_int, [_P, _P, C.c_int, C.c_int64, _P, C.c_int, C.c_int, _P, C.c_int64, C.c_int64, _P]),
    "zvx_symbols_create": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(_P)]),
    "zvx_symbols_destroy": (None, [_P]),
    "zvx_symbols_last_error": (C.c_char_p, [_P]),
    "zvx_symbols_num_phones": (C.c_int, [_P]),
    "zvx_symbols_num_puncts": (C.c_int, [_P]),
    "zvx_transcript2phonemids": (C.c_int, [_P, C.c_char_p, _P, _P, C.c_int]),
    "zvx_collate": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "zvx_ragged_pack": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, _P, _P, _P, _P]),
    "zvx_ragged_unpack": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int,
                                    _P]),
    "zvx_ragged_last_error": (C.c_char_p, []),
    "zvx_attention": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P]),
    "zvx_attention_ex": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, C.c_int, C.c_int, _P, C.c_int64, _P]),
    "zvx_attention_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "zvx_attention_last_error": (C.c_char_p, []),
    "zvx_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "zvx_workspace_bytes": (C.c_int64, [_P]),
    "zvx_launch_count": (C.c_int64, [_P]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the zerovox_b200 CUDA library is not built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc, sm_100a). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.zvx_abi_version() != ZVX_ABI_VERSION:
        raise ImportError("libzerovox_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib
