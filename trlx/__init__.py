"""Drop-in alias: ``import trlx`` (and every ``trlx.<sub>.<module>`` path of the reference) resolves to ``trlx_b200``."""
import importlib
import importlib.abc
import importlib.util
import sys

import trlx_b200 as _impl


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    prefix = "trlx."

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(self.prefix):
            return None
        real = "trlx_b200." + fullname[len(self.prefix):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, ValueError):
            return None
        return importlib.util.spec_from_loader(fullname, self, is_package=True)

    def create_module(self, spec):
        real = importlib.import_module("trlx_b200." + spec.name[len(self.prefix):])
        return real

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _AliasFinder())
train = _impl.train
logging = _impl.logging
__version__ = _impl.__version__
