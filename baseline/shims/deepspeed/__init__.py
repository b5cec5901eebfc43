import contextlib
import types


@contextlib.contextmanager
def _gathered(params, modifier_rank=None, **kw):
    yield


zero = types.SimpleNamespace(GatheredParameters=_gathered)
comm = types.SimpleNamespace(get_rank=lambda: 0)
