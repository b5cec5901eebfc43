"""trlx_b200 — a Blackwell-native RLHF framework with the capabilities and public API of CarperAI/trlX."""
from trlx_b200.utils import logging  # noqa: F401

__version__ = "0.1.0"


def train(*args, **kwargs):
    """See :func:`trlx_b200.trlx.train` (imported lazily so that ``import trlx_b200`` stays light)."""
    from trlx_b200.trlx import train as _train

    return _train(*args, **kwargs)
