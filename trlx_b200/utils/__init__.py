"""General utilities (parity: ``trlx/utils/__init__.py``).

Differences from the reference that matter on B200:

* :class:`Clock` can time on the *device* (CUDA events) — the reference's wall-clock timers never
  synchronise, so they measure launch time.
* ``get_optimizer_class`` resolves ``adam``/``adamw`` to :class:`trlx_b200.parallel.optim.FusedAdamW`,
  the partitioned multi-tensor optimizer whose update runs in one sm_100a kernel (and, under
  data parallelism, is fused with the gradient reduce-scatter).  The ``*_8bit_bnb`` names resolve to
  an in-repo block-quantised state optimizer instead of requiring ``bitsandbytes``.
"""
from __future__ import annotations

import importlib.util
import math
import os
import random
import subprocess
import time
from dataclasses import fields, is_dataclass
from enum import Enum
from numbers import Number
from typing import Any, Callable, Dict, Iterable, Iterator, Optional, Tuple

import numpy as np
import torch


_SOURCE_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def output_root() -> Optional[str]:
    """Where relative output paths (checkpoints, logs, exports) go.

    ``$TRLX_B200_OUT`` wins.  Otherwise the current directory, like the reference (``checkpoint_dir="ckpts"`` is
    cwd-relative there) — unless the current directory is inside this source checkout: artefacts written there travel
    with every repository snapshot (a 498 MB example checkpoint once pushed one over its size limit), so they are
    redirected to ``<tmp>/trlx_b200_out``.  Returns ``None`` for "use cwd"."""
    env = os.environ.get("TRLX_B200_OUT")
    if env:
        return env
    cwd = os.path.realpath(os.getcwd())
    root = os.path.realpath(_SOURCE_ROOT)
    in_checkout = os.path.exists(os.path.join(root, "__graft_entry__.py")) and (cwd == root or cwd.startswith(root + os.sep))
    if in_checkout:
        import tempfile

        return os.path.join(tempfile.gettempdir(), "trlx_b200_out")
    return None


def resolve_output_dir(path: Optional[str], for_read: bool = False) -> Optional[str]:
    """Map a user-supplied output directory through :func:`output_root`.  Absolute paths pass through; with
    ``for_read`` an existing cwd-relative path is preferred (loading a checkpoint someone put there on purpose)."""
    if path is None or os.path.isabs(path):
        return path
    if for_read and os.path.exists(path):
        return path
    root = output_root()
    if root is None:
        return path
    out = os.path.join(root, path)
    if not for_read and not getattr(resolve_output_dir, "_told", False):
        resolve_output_dir._told = True
        import logging as _logging

        _logging.getLogger(__name__).warning(f"relative output paths are written under {root} (set TRLX_B200_OUT to choose)")
    return out


def is_peft_available() -> bool:
    """External ``peft`` is never required: adapters are implemented in ``trlx_b200.models.peft``."""
    return importlib.util.find_spec("peft") is not None


def rank() -> int:
    return int(os.environ.get("RANK", "0"))


def world_size() -> int:
    return int(os.environ.get("WORLD_SIZE", "1"))


def local_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


def print_rank_0(*message) -> None:
    if rank() == 0:
        print(*message)


def significant(x: Any, ndigits: int = 2) -> Any:
    """Round to ``ndigits`` digits after the leading significant digit; non-numbers pass through."""
    if isinstance(x, torch.Tensor):
        x = x.item()
    if not isinstance(x, Number) or isinstance(x, bool):
        return x
    if x == 0 or math.isnan(x) or math.isinf(x):
        return x
    magnitude = int(math.floor(math.log10(abs(x))))
    return round(x, ndigits - magnitude)


def set_seed(seed: int, parallel=None) -> None:
    """Seed python / numpy / torch with ``seed + data-parallel rank``: DP replicas sample different rollouts, while the
    tensor/pipeline-parallel ranks of one replica share a stream (``rank = (pp·DP + dp)·TP + tp``, parallel/runtime.py)."""
    r = rank()
    if parallel is not None:
        tp = max(int(getattr(parallel, "tensor_parallel", 1) or 1), 1)
        pp = max(int(getattr(parallel, "pipeline_parallel", 1) or 1), 1)
        dp = max(int(os.environ.get("WORLD_SIZE", "1")) // (tp * pp), 1)
        r = (r // tp) % dp
    seed = int(seed) + r
    random.seed(seed)
    np.random.seed(seed % (2**32))
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def get_distributed_config(accelerator=None, runtime=None) -> Dict[str, Any]:
    """Summary of the parallel layout for trackers (reference: ``utils/__init__.py:58-80``, which takes an ``Accelerator``;
    here the first argument is the :class:`trlx_b200.parallel.runtime.Runtime` — anything with ``describe()``)."""
    runtime = runtime if runtime is not None else accelerator
    cfg = {"mixed_precision": "bf16", "num_gpus": world_size()}
    if runtime is not None and hasattr(runtime, "describe"):
        cfg.update(runtime.describe())
    return cfg


# ---- optimizers / schedulers -----------------------------------------------------------------
class OptimizerName(str, Enum):
    ADAM = "adam"
    ADAMW = "adamw"
    ADAM_8BIT_BNB = "adam_8bit_bnb"
    ADAMW_8BIT_BNB = "adamw_8bit_bnb"
    SGD = "sgd"
    DISTRIBUTED_FUSED_ADAM = "distributed_fused_adam"  # NeMo/Apex name: the sharded fused AdamW is exactly that here


def get_optimizer_class(name):
    """Optimizer class for ``name`` (str or :class:`OptimizerName`)."""
    key = name.value if isinstance(name, OptimizerName) else str(name)
    from trlx_b200.parallel import optim as _optim

    table = {
        OptimizerName.ADAM.value: _optim.FusedAdam,
        OptimizerName.ADAMW.value: _optim.FusedAdamW,
        OptimizerName.ADAM_8BIT_BNB.value: _optim.Adam8bit,
        OptimizerName.ADAMW_8BIT_BNB.value: _optim.AdamW8bit,
        OptimizerName.SGD.value: torch.optim.SGD,
        OptimizerName.DISTRIBUTED_FUSED_ADAM.value: _optim.FusedAdamW,
    }
    if key not in table:
        supported = [o.value for o in OptimizerName]
        raise ValueError(f"`{name}` is not a supported optimizer. Supported optimizers are: {supported}")
    return table[key]


class SchedulerName(str, Enum):
    COSINE_ANNEALING = "cosine_annealing"
    LINEAR = "linear"
    NEMO_COSINE_ANNEALING = "CosineAnnealing"  # NeMo's warmup → cosine → constant schedule (nemo_* examples)


class WarmupCosineAnnealing(torch.optim.lr_scheduler.LambdaLR):
    """``warmup_steps`` linear warm-up, cosine decay to ``min_lr`` until ``max_steps - constant_steps``, then constant
    (the schedule the reference's NeMo examples configure: ``examples/nemo_ppo_sentiments.py:69-72``)."""

    def __init__(self, optimizer, warmup_steps: int = 0, constant_steps: float = 0, min_lr: float = 0.0,
                 max_steps: float = 1e12, last_epoch: int = -1, **unused):
        import math

        base = [g["lr"] for g in optimizer.param_groups]
        decay_steps = max(float(max_steps) - float(constant_steps) - warmup_steps, 1.0)

        def factor(base_lr):
            def f(step):
                if warmup_steps and step < warmup_steps:
                    return (step + 1) / (warmup_steps + 1)
                t = min((step - warmup_steps) / decay_steps, 1.0)
                lr = min_lr + 0.5 * (base_lr - min_lr) * (1 + math.cos(math.pi * t))
                return lr / base_lr if base_lr else 1.0
            return f

        super().__init__(optimizer, [factor(b) for b in base], last_epoch=last_epoch)


def get_scheduler_class(name):
    key = name.value if isinstance(name, SchedulerName) else str(name)
    from torch.optim.lr_scheduler import CosineAnnealingLR, LinearLR

    table = {SchedulerName.COSINE_ANNEALING.value: CosineAnnealingLR, SchedulerName.LINEAR.value: LinearLR,
             SchedulerName.NEMO_COSINE_ANNEALING.value: WarmupCosineAnnealing}
    if key not in table:
        supported = [s.value for s in SchedulerName]
        raise ValueError(f"`{name}` is not a supported scheduler. Supported schedulers are: {supported}")
    return table[key]


# ---- timing ------------------------------------------------------------------------------------
class Clock:
    """Stopwatch.  ``tick`` returns seconds since the previous tick.

    With ``device=True`` on a CUDA machine, ticks are CUDA events recorded on the current
    stream and ``tick`` returns *device* time (it synchronises on the previous event only).
    """

    def __init__(self, device: bool = False):
        self._device = bool(device) and torch.cuda.is_available()
        self.total_time = 0.0
        self.total_samples = 0
        self._mark()

    def _mark(self):
        if self._device:
            self._ev = torch.cuda.Event(enable_timing=True)
            self._ev.record()
        self.start = time.time()

    def tick(self, samples: int = 0) -> float:
        if self._device:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            end.synchronize()
            delta = self._ev.elapsed_time(end) / 1e3
            self._ev = end
            self.start = time.time()
        else:
            now = time.time()
            delta = now - self.start
            self.start = now
        if samples:
            self.total_time += delta
            self.total_samples += samples
        return delta

    def get_stat(self, n_samp: int = 1000, reset: bool = False) -> float:
        """Average seconds per ``n_samp`` samples."""
        per = self.total_time / max(self.total_samples, 1)
        if reset:
            self.total_time, self.total_samples = 0.0, 0
        return per * n_samp


# ---- pytrees -----------------------------------------------------------------------------------
def tree_map(f: Callable, tree: Any) -> Any:
    """Apply ``f`` to every leaf of a nest of dataclasses / dicts / lists / tuples."""
    if is_dataclass(tree) and not isinstance(tree, type):
        return type(tree)(**{fl.name: tree_map(f, getattr(tree, fl.name)) for fl in fields(tree)})
    if isinstance(tree, dict):
        return {k: tree_map(f, v) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, v) for v in tree)
    if hasattr(tree, "data") and hasattr(tree, "keys") and not isinstance(tree, torch.Tensor):
        # BatchEncoding-like mapping
        return type(tree)({k: tree_map(f, v) for k, v in tree.items()})
    return f(tree)


def to_device(tree: Any, device, non_blocking: bool = False) -> Any:
    return tree_map(lambda x: x.to(device, non_blocking=non_blocking) if isinstance(x, torch.Tensor) else x, tree)


def filter_non_scalars(xs: Dict) -> Dict:
    """Keep only entries castable to ``float``."""
    out = {}
    for k, v in xs.items():
        try:
            out[k] = float(v)
        except (TypeError, ValueError):
            pass
    return out


def get_git_tag() -> Tuple[str, str]:
    """``(branch, 'hash/date')`` of HEAD, or ``('unknown','unknown')`` outside a work tree."""
    try:
        desc = subprocess.check_output(["git", "log", "--format=%h/%as", "-n1"], stderr=subprocess.DEVNULL)
        branch = subprocess.check_output(["git", "rev-parse", "--abbrev-ref", "HEAD"], stderr=subprocess.DEVNULL)
        return branch.decode().strip(), desc.decode().strip()
    except (subprocess.CalledProcessError, FileNotFoundError, OSError):
        return "unknown", "unknown"


def infinite_dataloader(dataloader: Iterable, sampler=None) -> Iterator:
    """Cycle over ``dataloader`` forever, bumping the (distributed) sampler's epoch each pass."""
    epoch = 0
    while True:
        if sampler is not None and hasattr(sampler, "set_epoch"):
            sampler.set_epoch(epoch)
        epoch += 1
        yielded = False
        for item in dataloader:
            yielded = True
            yield item
        if not yielded:
            raise ValueError("infinite_dataloader: underlying loader is empty")
