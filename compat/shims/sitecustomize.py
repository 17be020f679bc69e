"""Import hook for `torch._six` (util/misc.py:21 of the reference does `from torch._six import inf`; the module was removed in
torch 2.0).  Installed at interpreter start-up when this directory is on PYTHONPATH; torch itself is not imported here."""
import importlib.abc
import importlib.machinery
import sys
import types


class _TorchSix(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name == "torch._six":
            return importlib.machinery.ModuleSpec(name, self)
        return None

    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.inf = float("inf")
        m.string_classes = (str, bytes)
        return m

    def exec_module(self, module):
        pass


if not any(isinstance(f, _TorchSix) for f in sys.meta_path):
    sys.meta_path.append(_TorchSix())
