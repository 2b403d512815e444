"""Importable alias of the package directory ``gansformer-reproducibility-challenge_b200/`` (its name is not a
valid Python identifier).  ``import gansformer_b200 as gf`` gives the same module object."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("gansformer-reproducibility-challenge_b200")
sys.modules[__name__] = _pkg
