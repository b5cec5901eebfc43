"""Trainer registry and the backend-agnostic base class (parity: ``trlx/trainer/__init__.py``).

``@register_trainer`` files a class under its lower-cased name; :func:`trlx_b200.utils.loading.get_trainer` looks names up
case-insensitively (``"AcceleratePPOTrainer"``, ``"NeMoILQLTrainer"`` …), which is how ``train.trainer`` in a config selects
the implementation.
"""
from __future__ import annotations

from abc import abstractmethod
from typing import Any, Callable, Dict, Iterable, Optional

from trlx_b200.data.configs import TRLConfig
from trlx_b200.pipeline import BaseRolloutStore
from trlx_b200.utils.registry import Registry

_TRAINERS: Registry = Registry("trainer")


def register_trainer(name=None):
    """Class decorator (bare or with an explicit alias) that adds a trainer to the registry."""
    return _TRAINERS.register(name)


@register_trainer
class BaseRLTrainer:
    """State every trainer shares, independent of the runtime underneath.

    ``reward_fn(samples, prompts, outputs, **metadata)`` scores rollouts of the online methods, ``metric_fn`` (same signature,
    returns named lists) is evaluated on generations during evaluation, ``logit_mask`` restricts which token may follow which
    (ILQL), ``stop_sequences`` trim generations, ``train_mode`` tells subclasses whether optimizer state is needed (inference-only
    construction skips it).  ``store`` is created by the subclass and holds the experience :meth:`learn` consumes."""

    def __init__(self, config: TRLConfig, reward_fn: Optional[Callable] = None, metric_fn: Optional[Callable] = None,
                 logit_mask=None, stop_sequences=None, train_mode: bool = False):
        self.config = config
        self.reward_fn, self.metric_fn = reward_fn, metric_fn
        self.logit_mask, self.stop_sequences = logit_mask, stop_sequences
        self.train_mode = train_mode
        self.store: BaseRolloutStore = None

    def push_to_store(self, data: Iterable[Any]) -> None:
        """Append experience (elements of the method's datatype) to ``self.store``."""
        self.store.push(data)

    @abstractmethod
    def learn(self) -> Optional[Dict[str, Any]]:
        """Optimise the model on the contents of ``self.store`` (and, for online methods, refill it between epochs)."""
