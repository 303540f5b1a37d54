"""CPU: the C-ABI library loads and exports exactly what include/afan_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cv_a-fan_b200", "libafan_b200.so")
HEADER = os.path.join(ROOT, "include", "afan_b200.h")


def declared_functions():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(afan_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        import __graft_entry__ as ge
        ge.build()
    return LIB


def test_header_symbols_all_exported(built):
    lib = ctypes.CDLL(built)
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/afan_b200.h but not exported"
    out = subprocess.check_output(["nm", "-D", "--defined-only", built], text=True)
    exported = sorted(set(re.findall(r"\b(afan_[a-z0-9_]+)\b", out)))
    assert exported == names, "exported afan_* symbols and the header differ"


def test_python_binding_table_matches_header(built):
    import importlib
    pkg = importlib.import_module("cv_a-fan_b200")
    assert sorted(pkg._lib.SIGNATURES) == declared_functions()
    assert pkg._lib.lib().afan_version().startswith(b"afan_b200")
    assert pkg._lib.lib().afan_strerror(-3) == b"workspace missing, misaligned or too small"


def test_no_compute_without_gpu_but_sizes_and_errors_work(built):
    import importlib
    L = importlib.import_module("cv_a-fan_b200")._lib.lib()
    assert L.afan_bn_workspace_bytes(2, 64) > 0
    assert L.afan_bn_workspace_bytes(0, 64) < 0
    assert L.afan_pgd_norms_workspace_bytes(128) >= 128 * 4
    # argument validation happens before any CUDA call
    assert L.afan_pgd_linf_step_f32(None, None, None, None, None, None, 0, 4, 4, 0.1, 0.1, 1, None) == -1
    assert L.afan_pgd_linf_step_f32(None, None, None, None, None, None, 0, -1, 4, 0.1, 0.1, 1, None) == -2
    assert L.afan_mix_feature_f32(None, None, None, 1, 1, 1, None) == -1
    assert L.afan_sgd_momentum_f32(None, None, None, 0, None, 0.9, 0.0, 1.0, None) == 0


def test_product_package_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "cv_a-fan_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "libafan_oracle" not in txt, f


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import importlib
    lib_mod = importlib.import_module("cv_a-fan_b200")._lib
    monkeypatch.setattr(lib_mod, "_lib", None)
    monkeypatch.setattr(lib_mod, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(lib_mod.AfanError, match="no CPU/PyTorch fallback"):
        lib_mod.lib()
