"""CPU: the C-ABI library loads and exports every symbol include/zerovox_b200.h declares; the ctypes mirror of
zvx_config has the C layout; calls fail loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from zerovox_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "zerovox_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zvx_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"libzerovox_b200.so does not export {n}"
    assert set(names) == set(_lib.SYMBOLS), "ctypes table out of sync with the header"
    assert lib.zvx_abi_version() == _lib.ZVX_ABI_VERSION


def test_config_struct_layout_matches_c(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "zerovox_b200.h"\n'
                    'int main(){printf("%zu %zu %zu %zu\\n", sizeof(zvx_config), offsetof(zvx_config, hg_resblock),'
                    ' offsetof(zvx_config, hg_resblock_dilation_sizes), offsetof(zvx_config, tensor_core_policy));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    size, o1, o2, o3 = map(int, subprocess.check_output([str(exe)]).split())
    assert size == C.sizeof(_lib.ZvxConfig)
    assert o1 == _lib.ZvxConfig.hg_resblock.offset
    assert o2 == _lib.ZvxConfig.hg_resblock_dilation_sizes.offset
    assert o3 == _lib.ZvxConfig.tensor_core_policy.offset


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    from zerovox_b200.engine import EngineConfig
    lib = _lib.load()
    cfg = EngineConfig().to_c()
    h = C.c_void_p()
    rc = lib.zvx_create(C.byref(cfg), 0, C.byref(h))
    assert rc != 0 and not h
    assert lib.zvx_last_error(None)  # message present
    # null-handle calls are rejected, not crashed
    assert lib.zvx_finalize_weights(None) != 0
    assert lib.zvx_vocode(None, None, 1, 1, None, None) != 0


def test_bad_abi_version_rejected():
    from zerovox_b200.engine import EngineConfig
    lib = _lib.load()
    cfg = EngineConfig().to_c()
    cfg.abi_version = 999
    h = C.c_void_p()
    assert lib.zvx_create(C.byref(cfg), 0, C.byref(h)) != 0


def test_mel_config_struct_layout_matches_c(tmp_path):
    prog = tmp_path / "szm.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "zerovox_b200.h"\n'
                    'int main(){printf("%zu %zu %zu\\n", sizeof(zvx_mel_config), offsetof(zvx_mel_config, fmin),'
                    ' offsetof(zvx_mel_config, reserved));return 0;}\n')
    exe = tmp_path / "szm"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    size, o1, o2 = map(int, subprocess.check_output([str(exe)]).split())
    assert size == C.sizeof(_lib.ZvxMelConfig)
    assert o1 == _lib.ZvxMelConfig.fmin.offset and o2 == _lib.ZvxMelConfig.reserved.offset


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_frontend_fails_loudly_without_gpu():
    from zerovox_b200.frontend import MelFrontend
    lib = _lib.load()
    cfg = _lib.ZvxMelConfig()
    cfg.abi_version, cfg.sampling_rate, cfg.fft_size, cfg.hop_size, cfg.win_length, cfg.num_mels = 1, 22050, 1024, 256, 1024, 80
    cfg.fmin, cfg.fmax = 0.0, 8000.0
    h = C.c_void_p()
    assert lib.zvx_frontend_create(C.byref(cfg), 0, C.byref(h)) != 0 and not h      # no device: no handle, no CPU path
    assert lib.zvx_frontend_last_error(None)
    assert lib.zvx_mel_spectrogram(None, None, 1, 1, None, None, 1, None, None, None) != 0
    assert lib.zvx_trim_silence(None, None, 1, 1, None, 40.0, 2048, 512, None, None, None, None) != 0
    with pytest.raises(RuntimeError, match="no CPU path"):
        MelFrontend()


def test_tokeniser_abi_capacity_protocol():
    """zvx_transcript2phonemids writes at most `capacity` ids and returns the needed count (host code, runs anywhere)."""
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.zvx_symbols_create(b"abc", b" ,.", C.byref(h)) == 0
    ph, pu = (C.c_int32 * 2)(-7, -7), (C.c_int32 * 2)(-7, -7)
    assert lib.zvx_transcript2phonemids(h, b"ab, cab.", ph, pu, 2) == 5          # needs 5, wrote 2
    assert list(ph) == [0, 1] and list(pu) == [0, 2]
    ph5, pu5 = (C.c_int32 * 5)(), (C.c_int32 * 5)()
    assert lib.zvx_transcript2phonemids(h, b"ab, cab.", ph5, pu5, 5) == 5
    assert list(ph5) == [0, 1, 2, 0, 1] and list(pu5) == [0, 2, 0, 0, 3]
    assert lib.zvx_transcript2phonemids(h, b"", ph5, pu5, 5) == 0
    assert lib.zvx_transcript2phonemids(None, b"a", ph5, pu5, 5) < 0
    assert lib.zvx_symbols_num_phones(h) == 3 and lib.zvx_symbols_num_puncts(h) == 4
    lib.zvx_symbols_destroy(h)
