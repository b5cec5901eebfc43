"""Method-config registry (parity: ``trlx/data/method_configs.py``)."""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any, Dict

from trlx_b200.utils.registry import Registry

_METHODS: Registry = Registry("method config")


def register_method(name=None):
    """``@register_method`` / ``@register_method("name")`` — names are lower-cased."""
    return _METHODS.register(name)


@dataclass
class MethodConfig:
    """Base class of every RL-method hyper-parameter block.

    :param name: registered (case-insensitive) name of the concrete config class
    """

    name: str

    @classmethod
    def from_dict(cls, config: Dict[str, Any]):
        return cls(**config)

    def to_dict(self) -> Dict[str, Any]:
        return {f.name: getattr(self, f.name) for f in fields(self)}


register_method(MethodConfig)


def get_method(name: str):
    """Return the config class registered under ``name``."""
    try:
        return _METHODS.get(name)
    except KeyError:
        pass
    # PPOConfig / ILQLConfig / SFTConfig / RFTConfig register themselves when their modules are imported; a bare
    # ``TRLConfig.load_yaml`` may run before anything imported them
    import importlib

    importlib.import_module("trlx_b200.utils.loading")
    try:
        return _METHODS.get(name)
    except KeyError as e:
        raise Exception(f"Error: Trying to access a method that has not been registered ({e})") from None
