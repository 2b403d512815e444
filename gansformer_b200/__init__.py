"""Importable alias of the package directory ``gansformer-reproducibility-challenge_b200/`` (its name is not a
valid Python identifier).  ``import gansformer_b200 as gf`` gives the same module object, and every submodule
(``gansformer_b200.training`` ...) IS the submodule of the real package -- one copy of every module-level switch
(``networks.CACHE_BYPASS``, ``attention.STAGE_TIMER``, the ``_lib`` handle), one class object per class."""
import importlib
import importlib.abc
import importlib.util
import os
import sys

_REAL = "gansformer-reproducibility-challenge_b200"
_ALIAS = __name__

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name):
        self.real_name = real_name

    def create_module(self, spec):
        return importlib.import_module(self.real_name)      # the real module object, not a second copy

    def exec_module(self, module):                          # already executed under its real name
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    """``gansformer_b200.x.y`` -> the module object of ``gansformer-reproducibility-challenge_b200.x.y``."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != _ALIAS and not fullname.startswith(_ALIAS + "."):
            return None
        real = _REAL + fullname[len(_ALIAS):]
        try:
            real_spec = importlib.util.find_spec(real)
        except (ImportError, ValueError):
            return None
        if real_spec is None:
            return None
        return importlib.util.spec_from_loader(fullname, _AliasLoader(real), is_package=real_spec.submodule_search_locations is not None)


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):               # submodules the real package has already imported
    if _name.startswith(_REAL + "."):
        sys.modules[_ALIAS + _name[len(_REAL):]] = _mod
sys.modules[_ALIAS] = _pkg
