"""Python face of the sm_100a kernels in ``trlx_b200/csrc``.

* ``ops.C``            — the compiled extension module (``trlx_b200/_C.so``), loaded lazily.
* ``ops.available()``  — True when a CUDA device is present AND the extension loads.  On a GPU box a
  missing/broken extension is a hard error (no silent eager fallback) unless ``TRLX_B200_ALLOW_EAGER=1``.
* autograd wrappers (:func:`linear`, :func:`fused_logprob`, :func:`ppo_loss`, …) whose forward runs the
  hand-written kernels; every wrapper has a plain PyTorch twin in :mod:`trlx_b200.ops.reference` that is both the
  CPU path and the numerical oracle for the tests.
"""
from __future__ import annotations

import functools
import importlib
import os
from typing import Optional

import torch

from trlx_b200.ops import reference  # noqa: F401


class ExtensionMissing(RuntimeError):
    pass


@functools.lru_cache(maxsize=1)
def _load():
    try:
        return importlib.import_module("trlx_b200._C"), None
    except Exception as first:  # not built yet → try building in-tree once
        try:
            from trlx_b200.csrc.build import build

            build()
            return importlib.import_module("trlx_b200._C"), None
        except Exception as second:
            return None, f"{type(first).__name__}: {first}; build attempt: {type(second).__name__}: {second}"


def extension():
    mod, err = _load()
    if mod is None:
        raise ExtensionMissing(f"trlx_b200._C is not available ({err})")
    return mod


class _Lazy:
    def __getattr__(self, name):
        return getattr(extension(), name)


C = _Lazy()


@functools.lru_cache(maxsize=1)
def available() -> bool:
    """CUDA device present and extension importable.  Raises on a GPU box without the extension."""
    if not torch.cuda.is_available():
        return False
    mod, err = _load()
    if mod is None:
        if os.environ.get("TRLX_B200_ALLOW_EAGER", "0") == "1":
            return False
        raise ExtensionMissing(
            f"a CUDA device is present but the sm_100a extension failed to load ({err}); "
            "run `python -m trlx_b200.csrc.build` or set TRLX_B200_ALLOW_EAGER=1 to accept the eager fallback"
        )
    major, _minor = torch.cuda.get_device_capability()
    return major >= 10


def enabled_for(*tensors) -> bool:
    """Kernel path applies: extension available and every tensor is CUDA bf16."""
    if not tensors or not all(isinstance(t, torch.Tensor) and t.is_cuda for t in tensors):
        return False
    return available()


from trlx_b200.ops.functional import (  # noqa: E402,F401
    fused_logprob,
    gae_and_whiten,
    linear,
    logprobs_from_logits,
    ppo_loss,
)
