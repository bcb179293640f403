"""The C-ABI library loads and exports every symbol include/*.h declares (no GPU needed)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    names = set()
    for h in (ROOT / "include").glob("*.h"):
        names |= set(re.findall(r"\b(b9gw_\w+)\s*\(", h.read_text()))
    return names


def test_header_and_binding_declare_the_same_symbols():
    from base_b200 import groundwork as gw
    assert declared_symbols() == set(gw.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(str(built.CUDA_LIB))
    for name in sorted(declared_symbols()):
        assert hasattr(L, name), name


def test_library_is_sm_100a_only(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", str(built.CUDA_LIB)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_abi_version_and_loud_failure_without_a_device(built):
    from base_b200 import groundwork as gw
    assert gw.lib().b9gw_abi_version() == gw.ABI_VERSION
    if gw.device_count() > 0:
        pytest.skip("a CUDA device is visible; the no-device path cannot be exercised")
    import numpy as np
    for call in (lambda: gw.dfma_peak(0), lambda: gw.transcendental_rate("exp"),
                 lambda: gw.device_map("exp", np.zeros(4)), lambda: gw.lse_rows(np.zeros((2, 2)))):
        with pytest.raises(gw.GroundworkError) as ei:
            call()
        assert ei.value.code == -1 and "no CPU fallback" in str(ei.value)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from base_b200 import groundwork as gw
    monkeypatch.setattr(gw, "_lib", None)
    monkeypatch.setattr(gw, "LIB_PATH", tmp_path / "libb9_groundwork.so")
    with pytest.raises(FileNotFoundError, match="no CPU fallback"):
        gw.lib()


def test_product_package_never_touches_the_checker():
    for py in (ROOT / "base_b200").rglob("*.py"):
        text = py.read_text()
        if py.name == "build.py":
            continue  # building the checker is not using it
        assert "groundwork_ref" not in text and "tests._ref" not in text and "from tests" not in text, py
