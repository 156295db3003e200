"""Importable alias of the ``neural-pde-solver_b200`` package (a hyphen is not a valid
identifier): ``import npde_b200`` == ``importlib.import_module("neural-pde-solver_b200")``."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("neural-pde-solver_b200")
sys.modules[__name__] = _pkg
