"""Process-wide model-parallel state (the role of Apex ``parallel_state`` in the reference's NeMo models,
``trlx/models/modeling_nemo_ppo.py:18-27``): which tensor- / pipeline-parallel group this process belongs to.

:class:`trlx_b200.parallel.runtime.Runtime` publishes its groups here when it builds them, so building blocks that are
constructed without an explicit group (``ParallelLinear``, ``vocab_parallel_cross_entropy`` …) find the layout of the
running job; unit tests and stand-alone scripts may call :func:`set_model_parallel` themselves.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional


@dataclass
class ModelParallelState:
    tp_group: Optional[Any] = None   # torch.distributed group of the tensor-parallel peers (None = not sharded)
    tp_rank: int = 0
    tp_size: int = 1
    pp_group: Optional[Any] = None
    pp_rank: int = 0
    pp_size: int = 1
    dp_group: Optional[Any] = None
    dp_rank: int = 0
    dp_size: int = 1


_STATE = ModelParallelState()


def set_model_parallel(**fields) -> ModelParallelState:
    """Replace the published layout (unspecified fields fall back to the single-process defaults)."""
    global _STATE
    _STATE = ModelParallelState(**fields)
    return _STATE


def get_model_parallel() -> ModelParallelState:
    return _STATE


def get_tensor_model_parallel_world_size() -> int:
    return _STATE.tp_size


def get_tensor_model_parallel_rank() -> int:
    return _STATE.tp_rank


def get_tensor_model_parallel_group():
    return _STATE.tp_group


def get_pipeline_model_parallel_world_size() -> int:
    return _STATE.pp_size


def get_pipeline_model_parallel_rank() -> int:
    return _STATE.pp_rank


def get_data_parallel_world_size() -> int:
    return _STATE.dp_size


def get_data_parallel_rank() -> int:
    return _STATE.dp_rank
