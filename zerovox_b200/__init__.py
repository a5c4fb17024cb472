"""zerovox_b200 — B200-native (sm_100a) inference engine for ZeroVOX's phoneme -> mel -> waveform path.

Layout
  csrc/                 hand-written CUDA kernels + the C ABI (include/zerovox_b200.h)
  _lib.py, engine.py    ctypes binding and the tensor-level host API
  tts/                  mirrors of the reference's module interface (zerovox.tts.*): same constructor
                        arguments, forward()/inference_ex() signatures and state_dict keys
  patching.py           zerovox_b200.patch(): rebind only the eval-mode CUDA forward / inference_ex of the REFERENCE's
                        own ZeroVox class; training and CPU stay on the reference code
  parallel.py           batch sharding of utterances across GPUs (one scatter + one gather-v)
"""
from .engine import Engine, EngineConfig  # noqa: F401
from .patching import patch, unpatch  # noqa: F401

__version__ = "0.1.0"
