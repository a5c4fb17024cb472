"""zerovox_b200 — B200-native (sm_100a) inference engine for ZeroVOX's phoneme -> mel -> waveform path.

Layout
  csrc/                 hand-written CUDA kernels + the C ABI (include/zerovox_b200.h)
  _lib.py, engine.py    ctypes binding and the tensor-level host API
  tts/                  mirrors of the reference's module interface (zerovox.tts.*): same constructor
                        arguments, forward()/inference_ex() signatures and state_dict keys
  parallel.py           batch sharding of utterances across GPUs (one scatter + one gather)
"""
from .engine import Engine, EngineConfig  # noqa: F401

__version__ = "0.1.0"
