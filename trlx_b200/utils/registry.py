"""Case-insensitive name → class registries.

The reference keeps three hand-rolled module-level dicts with three copies of the
same decorator (``trlx/data/method_configs.py:9-28``, ``trlx/trainer/__init__.py:12-31``,
``trlx/pipeline/__init__.py:19-38``).  Here a single ``Registry`` object serves all
three and gives uniform error messages.
"""
from __future__ import annotations

from typing import Callable, Dict, Generic, Iterator, Optional, Type, TypeVar, Union

T = TypeVar("T")


class Registry(Generic[T]):
    def __init__(self, kind: str):
        self.kind = kind
        self._entries: Dict[str, Type[T]] = {}

    # ``@reg.register`` and ``@reg.register("alias")`` are both accepted.
    def register(self, target: Union[str, Type[T], None] = None) -> Union[Type[T], Callable[[Type[T]], Type[T]]]:
        if isinstance(target, str) or target is None:
            alias: Optional[str] = target

            def deco(cls: Type[T]) -> Type[T]:
                self._entries[(alias or cls.__name__).lower()] = cls
                return cls

            return deco
        self._entries[target.__name__.lower()] = target
        return target

    def get(self, name: str) -> Type[T]:
        try:
            return self._entries[name.lower()]
        except KeyError:
            known = ", ".join(sorted(self._entries)) or "<none>"
            raise KeyError(f"No {self.kind} registered under '{name}'. Known: {known}") from None

    def __contains__(self, name: str) -> bool:
        return name.lower() in self._entries

    def __iter__(self) -> Iterator[str]:
        return iter(self._entries)

    def items(self):
        return self._entries.items()
