"""Trainer registry and the backend-agnostic base class (parity: ``trlx/trainer/__init__.py``)."""
from __future__ import annotations

from abc import abstractmethod
from typing import Any, Callable, Dict, Iterable, Optional

from trlx_b200.data.configs import TRLConfig
from trlx_b200.pipeline import BaseRolloutStore
from trlx_b200.utils.registry import Registry

_TRAINERS: Registry = Registry("trainer")


def register_trainer(target=None):
    """Register a trainer class under its lower-cased name (or an explicit alias)."""
    return _TRAINERS.register(target)


@register_trainer
class BaseRLTrainer:
    def __init__(self, config: TRLConfig, reward_fn: Optional[Callable] = None, metric_fn: Optional[Callable] = None,
                 logit_mask=None, stop_sequences=None, train_mode: bool = False):
        self.store: BaseRolloutStore = None
        self.config = config
        self.reward_fn = reward_fn
        self.metric_fn = metric_fn
        self.logit_mask = logit_mask
        self.train_mode = train_mode
        self.stop_sequences = stop_sequences

    def push_to_store(self, data):
        """Append new experience to the rollout store."""
        self.store.push(data)

    @abstractmethod
    def learn(self):
        """Consume the rollout store to update the model."""
