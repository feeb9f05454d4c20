"""CPU: the C-ABI library loads and exports every symbol include/*.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(upk_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_declares_symbols():
    d = _declared()
    assert "upk_furthest_point_sampling" in d and "upk_ball_query" in d and len(d) >= 12


def test_library_exports_every_declared_symbol():
    from unopose_b200 import _lib, build

    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.upk_abi_version() >= 1
    assert lib.upk_built_sm() == 100


def test_python_binding_covers_header():
    from unopose_b200 import _lib

    assert set(_declared()) == set(_lib.exported_symbols())
    _lib.load()  # argtypes for every symbol resolve


def test_no_fallback_when_library_missing(monkeypatch):
    from unopose_b200 import _lib
    import pytest

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libunopose_b200.so")
    with pytest.raises(_lib.UnoposeNativeError):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "unopose_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)
