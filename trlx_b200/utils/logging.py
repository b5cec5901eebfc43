"""Library logger with rank filtering and a switchable progress bar.

Public surface follows ``trlx/utils/logging.py`` (env ``TRLX_VERBOSITY`` ``:47-60``; ``ranks=[...]``
kwarg and ``[RANK n]`` prefix ``:105-124``; verbosity / handler / propagation / format toggles
``:145-262``; ``TRLX_NO_ADVISORY_WARNINGS`` ``:264-272``; tqdm wrapper ``:278-340``).  State lives
in one ``_State`` object guarded by a lock rather than in module globals.
"""
from __future__ import annotations

import logging
import os
import sys
import threading
from logging import CRITICAL, DEBUG, ERROR, FATAL, INFO, NOTSET, WARN, WARNING  # noqa: F401
from typing import Dict, Optional

from tqdm import auto as _tqdm_auto

log_levels: Dict[str, int] = {
    "debug": logging.DEBUG,
    "info": logging.INFO,
    "warning": logging.WARNING,
    "error": logging.ERROR,
    "critical": logging.CRITICAL,
}
_DEFAULT_LEVEL = logging.INFO
_ROOT_NAME = __name__.split(".")[0]


class _State:
    lock = threading.Lock()
    handler: Optional[logging.Handler] = None
    progress_bars: bool = True


def _level_from_env() -> int:
    raw = os.getenv("TRLX_VERBOSITY")
    if raw is None:
        return _DEFAULT_LEVEL
    level = log_levels.get(raw.lower())
    if level is None:
        logging.getLogger().warning(
            f"Unknown option TRLX_VERBOSITY={raw}, has to be one of: {', '.join(log_levels)}"
        )
        return _DEFAULT_LEVEL
    return level


def _root() -> logging.Logger:
    return logging.getLogger(_ROOT_NAME)


def _ensure_configured() -> None:
    with _State.lock:
        if _State.handler is not None:
            return
        _State.handler = logging.StreamHandler(sys.stderr)
        _State.handler.flush = sys.stderr.flush  # type: ignore[method-assign]
        root = _root()
        root.addHandler(_State.handler)
        root.setLevel(_level_from_env())
        root.propagate = False


def _reset_library_root_logger() -> None:
    with _State.lock:
        if _State.handler is None:
            return
        root = _root()
        root.removeHandler(_State.handler)
        root.setLevel(logging.NOTSET)
        _State.handler = None


def get_log_levels_dict() -> Dict[str, int]:
    return log_levels


def _current_rank() -> int:
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank()
    except Exception:  # pragma: no cover - torch absent / half-initialised
        pass
    return int(os.environ.get("RANK", "0"))


class MultiProcessAdapter(logging.LoggerAdapter):
    """``logger.info(msg, ranks=["0", "3"])`` logs on those ranks only; ``ranks=[]`` = every rank.
    Default is rank 0 only.  Messages are prefixed with ``[RANK n]``."""

    def log(self, level, msg, *args, **kwargs):
        ranks = kwargs.pop("ranks", ["0"])
        mine = os.environ.get("RANK", "0")
        if not self.isEnabledFor(level):
            return
        if len(ranks) and mine not in {str(r) for r in ranks}:
            return
        msg, kwargs = self.process(msg, kwargs)
        self.logger._log(level, msg, args, **kwargs)

    def process(self, msg, kwargs):
        return f"[RANK {_current_rank()}] {msg}", kwargs

    def warning_advice(self, *args, **kwargs):
        """``warning`` that is silenced by ``TRLX_NO_ADVISORY_WARNINGS``."""
        if os.getenv("TRLX_NO_ADVISORY_WARNINGS", False):
            return
        self.warning(*args, **kwargs)


def get_logger(name: Optional[str] = None) -> MultiProcessAdapter:
    _ensure_configured()
    return MultiProcessAdapter(logging.getLogger(name or _ROOT_NAME), {})


def get_verbosity() -> int:
    _ensure_configured()
    return _root().getEffectiveLevel()


def set_verbosity(verbosity: int) -> None:
    _ensure_configured()
    _root().setLevel(verbosity)


def disable_default_handler() -> None:
    _ensure_configured()
    _root().removeHandler(_State.handler)


def enable_default_handler() -> None:
    _ensure_configured()
    if _State.handler not in _root().handlers:
        _root().addHandler(_State.handler)


def add_handler(handler: logging.Handler) -> None:
    _ensure_configured()
    assert handler is not None
    _root().addHandler(handler)


def remove_handler(handler: logging.Handler) -> None:
    _ensure_configured()
    assert handler is not None and handler in _root().handlers
    _root().removeHandler(handler)


def disable_propagation() -> None:
    _ensure_configured()
    _root().propagate = False


def enable_propagation() -> None:
    _ensure_configured()
    _root().propagate = True


def enable_explicit_format() -> None:
    """``[LEVEL|file:line] time >> message`` on every handler of the library logger."""
    _ensure_configured()
    fmt = logging.Formatter("[%(levelname)s|%(filename)s:%(lineno)s] %(asctime)s >> %(message)s")
    for h in _root().handlers:
        h.setFormatter(fmt)


def reset_format() -> None:
    _ensure_configured()
    for h in _root().handlers:
        h.setFormatter(None)


# ---- progress bars ------------------------------------------------------------------------------
class EmptyTqdm:
    """No-op stand-in used when progress bars are disabled."""

    def __init__(self, *args, **kwargs):
        self._iterator = args[0] if args else None

    def __iter__(self):
        return iter(self._iterator if self._iterator is not None else ())

    def __getattr__(self, _):
        return lambda *a, **k: None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _TqdmFactory:
    def __call__(self, *args, **kwargs):
        if _State.progress_bars:
            return _tqdm_auto.tqdm(*args, **kwargs)
        return EmptyTqdm(*args, **kwargs)

    def set_lock(self, *args, **kwargs):
        if _State.progress_bars:
            return _tqdm_auto.tqdm.set_lock(*args, **kwargs)

    def get_lock(self):
        if _State.progress_bars:
            return _tqdm_auto.tqdm.get_lock()


tqdm = _TqdmFactory()


def is_progress_bar_enabled() -> bool:
    return bool(_State.progress_bars)


def enable_progress_bar() -> None:
    _State.progress_bars = True


def disable_progress_bar() -> None:
    _State.progress_bars = False


def warning_advice(self, *args, **kwargs):
    """``logger.warning`` that stays silent when ``TRLX_NO_ADVISORY_WARNINGS`` is set (reference ``trlx/utils/logging.py:264-275``:
    a module-level function installed on ``logging.Logger``; the per-process adapter has the same method)."""
    if os.getenv("TRLX_NO_ADVISORY_WARNINGS", False):
        return
    self.warning(*args, **kwargs)


logging.Logger.warning_advice = warning_advice
