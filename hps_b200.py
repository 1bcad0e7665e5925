"""Importable alias of the package directory `hierarchical-point-slam_b200/` (a hyphen is not a valid
Python identifier): `import hps_b200` gives the very same module object."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("hierarchical-point-slam_b200")
