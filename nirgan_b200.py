"""Import alias: the package directory is ``nir-gan_b200`` (not a valid Python identifier), so
``import nirgan_b200`` resolves to it through this shim."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("nir-gan_b200")
sys.modules[__name__] = _pkg
