"""Importable alias of the `cv_a-fan_b200` package (its directory name is not a Python identifier).

`import afan_b200` and `python -m afan_b200.main_perturb` must resolve to the SAME module objects as
`importlib.import_module("cv_a-fan_b200")`: a second copy of `trainer` / `dual_bn` would make every
`isinstance(m, DualBatchNorm2d)` in the trainer fail silently.  So the alias covers the package AND every submodule,
including ones imported later (a meta-path finder maps `afan_b200.X` to `cv_a-fan_b200.X`).
"""
import importlib
import importlib.abc
import importlib.util
import sys

_REAL = "cv_a-fan_b200"
_ALIAS = __name__


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name):
        self.real_name = real_name

    def create_module(self, spec):
        return importlib.import_module(self.real_name)      # the one and only module object

    def exec_module(self, module):                          # already executed under its real name
        pass

    def get_code(self, fullname):                           # `python -m afan_b200.main_perturb` (runpy) asks for this
        return importlib.util.find_spec(self.real_name).loader.get_code(self.real_name)

    def is_package(self, fullname):
        return importlib.util.find_spec(self.real_name).submodule_search_locations is not None


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != _ALIAS and not fullname.startswith(_ALIAS + "."):
            return None
        real = _REAL + fullname[len(_ALIAS):]
        real_spec = importlib.util.find_spec(real)
        if real_spec is None:
            return None
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(real), origin=real_spec.origin)
        spec.has_location = real_spec.has_location
        return spec


_pkg = importlib.import_module(_REAL)
if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
sys.modules[_ALIAS] = _pkg
