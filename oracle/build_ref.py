#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: the unmodified reference modules of the path, copied from where they lie.

Runs ONLY where /root/reference exists (the build container); ``__graft_entry__.build()`` calls it.  The copy is
git-ignored (the reference's sources never enter this repository's history) but NOT gpurun-ignored, so the GPU box can
time the reference's own PyTorch CPU implementation (``bench.py --impl reference``, ``cpu_baseline.kind = "reference"``)
and the GPU parity tests can compare against the real modules.  Nothing is compiled: the reference is pure Python.

Usage:  python oracle/build_ref.py
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.reference_modules import REF_COPY, REF_FILES  # noqa: E402

LIGHTNING_STUB = '''"""Stand-in for the `lightning` package (absent from the image): the two base classes zerovox/tts/model.py names."""
import torch.nn as nn


class LightningModule(nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass


class LightningDataModule:
    pass
'''


def build_ref(src_root: str = "/root/reference", verbose: bool = True) -> bool:
    src = os.path.join(src_root, "zerovox", "tts")
    if not os.path.isdir(src):
        if verbose:
            print(f"[build_ref] {src_root} not present: keeping whatever oracle/_ref already holds")
        return os.path.exists(os.path.join(REF_COPY, "zerovox", "tts", "model.py"))
    dst = os.path.join(REF_COPY, "zerovox", "tts")
    os.makedirs(dst, exist_ok=True)
    for pkg in (os.path.join(REF_COPY, "zerovox"), dst):
        open(os.path.join(pkg, "__init__.py"), "a").close()
    for f in REF_FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    os.makedirs(os.path.join(REF_COPY, "lightning"), exist_ok=True)
    with open(os.path.join(REF_COPY, "lightning", "__init__.py"), "w") as fh:
        fh.write(LIGHTNING_STUB)
    if verbose:
        print(f"[build_ref] {len(REF_FILES)} reference modules -> {dst}")
    return True


if __name__ == "__main__":
    build_ref()
