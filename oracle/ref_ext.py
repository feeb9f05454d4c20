"""ORACLE — test infrastructure only.

Loader for the reference's own pointnet2 extension, compiled UNMODIFIED by
oracle/build_ref_ext.py into oracle/_ref/ref_pointnet2_ext.so.  Returns the
pybind module (the 9 functions of _ext_src/src/bindings.cpp:11-24) or None
when the prebuilt file is absent.  It executes only on a GPU
("CPU not supported", sampling.cpp:39).
"""
import importlib.machinery
import importlib.util
import os

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ref_pointnet2_ext.so")
_mod = None


def available():
    return os.path.exists(_SO)


def load():
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        return None
    import torch  # noqa: F401  (libtorch symbols must be loaded first)

    loader = importlib.machinery.ExtensionFileLoader("ref_pointnet2_ext", _SO)
    spec = importlib.util.spec_from_loader("ref_pointnet2_ext", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _mod = mod
    return mod
