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


# kernels launched per binding call (used for the launch counter that bench.py reports as ``gpu_launches``)
_KERNELS_PER_CALL = {"lmhead": 2, "ppo_loss": 3, "gae": 2}
_launches = 0


def reset_launch_count() -> None:
    global _launches
    _launches = 0


def launch_count() -> int:
    """Number of hand-written kernels launched (graph replays included) since :func:`reset_launch_count`."""
    return _launches


def add_launches(n: int) -> None:
    global _launches
    _launches += int(n)


class _Lazy:
    """Attribute proxy for the extension; kernel entry points are wrapped to count launches."""

    def __init__(self):
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is not None:
            return fn
        raw = getattr(extension(), name)
        if not callable(raw) or isinstance(raw, type) or name in ("ppo_loss_num_outputs",):
            return raw
        per_call = _KERNELS_PER_CALL.get(name, 1)

        def counted(*a, __raw=raw, __n=per_call, **k):
            global _launches
            _launches += __n
            return __raw(*a, **k)

        counted.__name__ = name
        self._cache[name] = counted
        return counted


C = _Lazy()


@functools.lru_cache(maxsize=1)
def available() -> bool:
    """CUDA device present and extension importable.  Raises on a GPU box without the extension."""
    if not torch.cuda.is_available() or os.environ.get("TRLX_B200_DISABLE_KERNELS", "0") == "1":
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
    attention,
    fused_logprob,
    gae_and_whiten,
    layer_norm,
    linear,
    logprobs_from_logits,
    norm_ok,
    ppo_loss,
)
