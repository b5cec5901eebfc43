"""Pipeline / rollout-store base classes and the micro-batch splitter.

Parity: ``trlx/pipeline/__init__.py`` (registry ``:14-38``, ``BasePipeline`` ``:41-67``,
``BaseRolloutStore`` ``:70-102``, ``MiniBatchIterator`` ``:105-177``).
"""
from __future__ import annotations

from abc import abstractmethod
from collections.abc import Mapping
from dataclasses import fields, is_dataclass
from typing import Any, Callable, Iterable, List, Optional

from torch.utils.data import DataLoader, Dataset

from trlx_b200.data import GeneralElement, RLElement
from trlx_b200.utils import logging
from trlx_b200.utils.registry import Registry

logger = logging.get_logger(__name__)

_DATAPIPELINE: Registry = Registry("pipeline")


def register_datapipeline(name=None):
    """Register a pipeline class under its lower-cased name (or an explicit alias)."""
    return _DATAPIPELINE.register(name)


@register_datapipeline
class BasePipeline(Dataset):
    """Dataset of prompts / samples with a ``create_loader`` factory."""

    def __init__(self, path: str = "dataset"):
        super().__init__()

    @abstractmethod
    def __getitem__(self, index: int) -> GeneralElement:
        ...

    @abstractmethod
    def __len__(self) -> int:
        ...

    @abstractmethod
    def create_loader(self, batch_size: int, shuffle: bool, prep_fn: Optional[Callable] = None,
                      num_workers: int = 0) -> DataLoader:
        ...


class BaseRolloutStore(Dataset):
    """Experience container; ``history`` holds the elements."""

    def __init__(self, capacity: int = -1):
        self.history: Iterable[Any] = None
        self.capacity = capacity

    @abstractmethod
    def push(self, exps: Iterable[Any]):
        ...

    def __getitem__(self, index: int) -> RLElement:
        return self.history[index]

    def __len__(self) -> int:
        return len(self.history)

    @abstractmethod
    def create_loader(self, batch_size: int, shuffle: bool, prep_fn: Optional[Callable] = None,
                      num_workers: int = 0) -> DataLoader:
        ...


def _columns(batch) -> dict:
    if is_dataclass(batch):
        return {f.name: getattr(batch, f.name) for f in fields(batch)}
    if isinstance(batch, Mapping):
        return dict(batch.items())
    raise TypeError(f"MiniBatchIterator cannot slice a {type(batch).__name__}")


def _rebuild(batch, cols: dict):
    if is_dataclass(batch):
        out = type(batch)(**cols)
        for k, v in vars(batch).items():  # per-batch annotations that are not dataclass fields (e.g. ``width``)
            if k not in cols:
                setattr(out, k, v)
        return out
    try:
        return type(batch)(cols)  # BatchEncoding / dict
    except Exception:
        return cols


class MiniBatchIterator:
    """Yields, for every batch of ``data_loader``, a list of ≤ ``num_mb`` micro-batches of
    ``mb_size`` rows each (gradient-accumulation slices).  Short tails produce a warning and a
    short/omitted micro-batch; an empty result ends iteration."""

    def __init__(self, data_loader, mb_size: int, num_mb: int):
        self.data_loader = data_loader
        self.data_loader_iter = iter(data_loader)
        self.mb_size = mb_size
        self.num_mb = num_mb

    def __iter__(self):
        return self

    def __next__(self) -> List[Any]:
        batch = next(self.data_loader_iter)
        if batch is None:
            logger.warning("MiniBatchIterator: the loader produced no batch — there are fewer samples than one minibatch needs; "
                           "add prompts / samples or lower `train.minibatch_size`")
            raise StopIteration
        cols = _columns(batch)
        n_total = min((len(v) for v in cols.values() if v is not None), default=0)
        if self.num_mb == 1 and 0 < n_total <= self.mb_size:
            return [batch]  # nothing to slice: hand the loader's batch through untouched
        out = []
        for i in range(self.num_mb):
            lo, hi = i * self.mb_size, (i + 1) * self.mb_size
            piece = {k: (v[lo:hi] if v is not None else None) for k, v in cols.items()}
            n_rows = min((len(v) for v in piece.values() if v is not None), default=0)
            if n_rows == 0:
                if self.num_mb > 1:
                    logger.warning(f"MiniBatchIterator: micro-batch {i} of {self.num_mb} is empty (batch of {n_total} rows, "
                                   f"mb_size {self.mb_size}): short last batch, or mb_size x num_mb exceeds the batch size")
                break
            if self.num_mb > 1 and n_rows < self.mb_size:
                logger.warning(f"MiniBatchIterator: micro-batch {i} holds {n_rows} of {self.mb_size} rows (short last batch, or "
                               "mb_size x num_mb exceeds the batch size)")
            out.append(_rebuild(batch, piece))
        if not out:
            raise StopIteration
        return out
