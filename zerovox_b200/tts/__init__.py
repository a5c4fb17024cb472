"""Mirrors of the reference's ``zerovox.tts`` module interface, backed by the CUDA engine."""
from .symbols import Symbols  # noqa: F401
from .model import ZeroVox, get_meldec, AttrDict  # noqa: F401
from .fs2 import FS2Encoder, FS2Decoder  # noqa: F401
from .hifigan import Generator  # noqa: F401
from .ResNetSE34V2 import ResNetSE34V2  # noqa: F401
from .styletts import StyleTTSDecoder  # noqa: F401
