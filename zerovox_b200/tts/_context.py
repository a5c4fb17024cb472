"""Engine ownership for trees of mirror modules.

Each mirror module (FS2Encoder, FS2Decoder, ResNetSE34V2, Generator) is a parameter container with the
reference's state_dict keys.  A context collects the parameters of the modules attached to it, pushes them
through zvx_set_weight under the reference's top-level prefixes and owns the engine handle.  A module used on
its own (e.g. the Generator returned by get_meldec) gets a private context; ZeroVox re-attaches its four
children to one shared context so a single handle (one workspace, one weight copy) serves the whole forward.
"""
from __future__ import annotations

import torch

from ..engine import Engine, EngineConfig

PREFIXES = {"encoder": "_phoneme_encoder.", "decoder": "_mel_decoder.", "spkemb": "_spkemb.", "vocoder": "_meldec."}


class EngineContext:
    def __init__(self):
        self.modules = {}      # role -> module
        self.engine = None
        self.stale = True
        self.tensor_core_policy = 1

    def attach(self, role: str, module):
        assert role in PREFIXES
        self.modules[role] = module
        object.__setattr__(module, "_ctx", self)
        self.engine = None
        self.stale = True

    def mark_stale(self):
        self.stale = True

    def config(self) -> EngineConfig:
        cfg = EngineConfig(tensor_core_policy=self.tensor_core_policy)
        for m in self.modules.values():
            m._fill_config(cfg)
        return cfg

    def get(self, device: torch.device) -> Engine:
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(
                "zerovox_b200 runs eval-mode forward on CUDA (sm_100a) only; move the module and its inputs to a "
                "CUDA device.  There is no CPU fallback (use the reference for CPU inference or training).")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self.engine is None or self.engine.device != device:
            self.engine = Engine(self.config(), device)
            self.stale = True
        if self.stale:
            for role, m in self.modules.items():
                self.engine.set_weights(m._engine_state_dict(), prefix=PREFIXES[role])
            self.engine.finalize()
            self.stale = False
        return self.engine


class EngineModuleMixin:
    """Shared behaviour of the mirror modules: staleness tracking + device lookup."""
    _role = None

    def _init_engine_binding(self):
        EngineContext().attach(self._role, self)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._ctx.mark_stale())

    def _apply(self, fn, *a, **k):  # .to() / .cuda() / .float() move the parameters -> re-push
        r = super()._apply(fn, *a, **k)
        ctx = getattr(self, "_ctx", None)
        if ctx is not None:
            ctx.mark_stale()
        return r

    def sync_weights(self):
        """Call after editing parameters in place."""
        self._ctx.mark_stale()

    def _device(self) -> torch.device:
        return next(self.parameters()).device

    def _engine(self) -> Engine:
        if self.training:
            raise NotImplementedError(
                f"{type(self).__name__}: the training-mode forward (autograd) is outside the zerovox_b200 hot path; "
                "call .eval() for inference or train with the reference implementation")
        return self._ctx.get(self._device())

    def _engine_state_dict(self):
        return self.state_dict()
