"""Importable alias of the `cv_a-fan_b200` package (its directory name is not a Python identifier)."""
import importlib
import sys

_pkg = importlib.import_module("cv_a-fan_b200")
sys.modules[__name__] = _pkg
