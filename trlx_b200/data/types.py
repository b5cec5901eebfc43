"""Element / batch records shared by pipelines, stores and trainers.

One module instead of the reference's four (``trlx/data/__init__.py``,
``accelerate_base_datatypes.py``, ``ppo_types.py:6-63``, ``ilql_types.py:1-139``); the
per-name modules re-export from here.  All records derive from :class:`Record`, which
adds device movement, dict export and the ``flatten_dataclass`` / ``unflatten_dataclass``
helpers that the reference's NeMo backend imports but never defined.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any, Callable, Iterable, List, Optional, Type

from torch import Tensor


class Record:
    """Mixin for tensor-holding dataclasses."""

    def to(self, device, non_blocking: bool = True):
        kw = {f.name: (v.to(device, non_blocking=non_blocking) if isinstance(v, Tensor) else v)
              for f in fields(self) for v in [getattr(self, f.name)]}  # type: ignore[arg-type]
        return type(self)(**kw)

    def pin(self):
        kw = {f.name: (v.pin_memory() if isinstance(v, Tensor) and not v.is_cuda else v)
              for f in fields(self) for v in [getattr(self, f.name)]}  # type: ignore[arg-type]
        return type(self)(**kw)

    def asdict(self):
        return {f.name: getattr(self, f.name) for f in fields(self)}  # type: ignore[arg-type]

    def __len__(self):  # batch records: leading dim of the first tensor field
        for f in fields(self):  # type: ignore[arg-type]
            v = getattr(self, f.name)
            if isinstance(v, Tensor):
                return v.shape[0]
            if isinstance(v, (list, tuple)):
                return len(v)
        return 0


def flatten_dataclass(cls: Type) -> Callable[[Any], List[Tensor]]:
    """``flatten_dataclass(PPORLBatch)(batch) -> [tensor, ...]`` in field order — lets a
    pipeline-parallel engine slice a batch into micro-batches tensor-by-tensor."""
    names = [f.name for f in fields(cls)]
    return lambda rec: [getattr(rec, n) for n in names]


def unflatten_dataclass(cls: Type) -> Callable[[Iterable[Tensor]], Any]:
    names = [f.name for f in fields(cls)]
    return lambda tensors: cls(**dict(zip(names, tensors)))


# ---- generic (vestigial in the reference, kept for import parity) -------------------------
@dataclass
class GeneralElement(Record):
    """General element produced by a data pipeline."""


@dataclass
class RLElement(Record):
    state: Optional[Iterable[str]] = None
    action: Optional[Tensor] = None
    reward: Optional[float] = None


@dataclass
class BatchElement(Record):
    tokens: Tensor
    masks: Tensor


@dataclass
class PromptElement(Record):
    text: str
    tokens: Tensor


@dataclass
class PromptBatch(Record):
    text: Iterable[str]
    tokens: Tensor


@dataclass
class AccelerateRLElement(Record):
    output_tokens: Tensor
    rewards: Tensor


@dataclass
class AccelerateRLBatchElement(Record):
    output_tokens: Tensor
    rewards: Tensor


# ---- PPO ----------------------------------------------------------------------------------
@dataclass
class PPORLElement(Record):
    """One rollout.

    :param query_tensor: prompt token ids ``[Q]``
    :param response_tensor: sampled token ids ``[R]``
    :param logprobs: behaviour log-probs of the response tokens ``[R]``
    :param values: value estimates at the response positions ``[R]``
    :param rewards: per-token rewards (KL penalty + score) ``[R]``
    """

    query_tensor: Tensor
    response_tensor: Tensor
    logprobs: Tensor
    values: Tensor
    rewards: Tensor


@dataclass
class PPORLBatch(Record):
    """Batched rollouts: queries left-padded ``[B,Q]``, everything else right-padded ``[B,R]``."""

    query_tensors: Tensor
    response_tensors: Tensor
    logprobs: Tensor
    values: Tensor
    rewards: Tensor


# ---- ILQL ---------------------------------------------------------------------------------
@dataclass
class ILQLElement(Record):
    input_ids: Tensor
    attention_mask: Tensor
    rewards: Tensor
    states_ixs: Tensor
    actions_ixs: Tensor
    dones: Tensor


@dataclass
class ILQLSeq2SeqElement(Record):
    input_ids: Tensor
    attention_mask: Tensor
    decoder_input_ids: Tensor
    rewards: Tensor
    states_ixs: Tensor
    actions_ixs: Tensor
    dones: Tensor


@dataclass
class ILQLBatch(Record):
    input_ids: Tensor
    attention_mask: Tensor
    rewards: Tensor
    states_ixs: Tensor
    actions_ixs: Tensor
    dones: Tensor


@dataclass
class ILQLSeq2SeqBatch(Record):
    input_ids: Tensor
    attention_mask: Tensor
    decoder_input_ids: Tensor
    rewards: Tensor
    states_ixs: Tensor
    actions_ixs: Tensor
    dones: Tensor
