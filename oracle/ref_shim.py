"""Import shim for the UNMODIFIED reference (VITA-Group/CV_A-FAN) on a CPU-only box.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py (to produce tests/golden/*.npz)
and by tests that cross-check the oracle restatement against the live reference when
/root/reference is mounted (this container; it does NOT exist on the GPU box).

The reference hard-codes `.cuda()` (Classification/attack_algo.py:44-46,
main_perturb.py:170-171) and imports `advertorch` (resnet_s.py:28) and matplotlib
(main_perturb.py:8), none of which are usable here.  The shim makes `.cuda()` the
identity and provides stand-ins for those two imports; no reference file is edited or
copied.
"""
import contextlib
import importlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("AFAN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Classification", "attack_algo.py"))


class _NormalizeByChannelMeanStd(nn.Module):
    """Stand-in for advertorch.utils.NormalizeByChannelMeanStd (used at resnet_s.py:87)."""

    def __init__(self, mean, std):
        super().__init__()
        self.register_buffer("mean", torch.as_tensor(mean, dtype=torch.float32))
        self.register_buffer("std", torch.as_tensor(std, dtype=torch.float32))

    def forward(self, x):
        return (x - self.mean[None, :, None, None]) / self.std[None, :, None, None]


def _install_stubs():
    adv = types.ModuleType("advertorch")
    adv_utils = types.ModuleType("advertorch.utils")
    adv_utils.NormalizeByChannelMeanStd = _NormalizeByChannelMeanStd
    adv.utils = adv_utils
    sys.modules.setdefault("advertorch", adv)
    sys.modules.setdefault("advertorch.utils", adv_utils)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            for name in ("plot", "legend", "savefig", "close", "figure", "imshow", "show", "subplot"):
                setattr(plt, name, lambda *a, **k: None)
            mpl.pyplot = plt
            mpl.use = lambda *a, **k: None
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt


@contextlib.contextmanager
def cpu_cuda_identity():
    """Make Tensor.cuda / Module.cuda the identity while reference code runs."""
    t_cuda, m_cuda = torch.Tensor.cuda, nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, nn.Module.cuda = t_cuda, m_cuda


def load(task: str, module: str):
    """Import `<REFERENCE_ROOT>/<task>/<module>.py` under a private name (e.g.
    load('Classification', 'attack_algo') -> module 'afan_ref.Classification.attack_algo')."""
    if not available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    task_dir = os.path.join(REFERENCE_ROOT, task)
    name = f"afan_ref_{task}_{module}"
    if name in sys.modules:
        return sys.modules[name]
    sys.path.insert(0, task_dir)
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(task_dir, module + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        sys.dont_write_bytecode, old = True, sys.dont_write_bytecode
        try:
            spec.loader.exec_module(mod)
        finally:
            sys.dont_write_bytecode = old
    finally:
        sys.path.remove(task_dir)
    return mod
