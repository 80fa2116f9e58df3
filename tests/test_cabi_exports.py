"""The C-ABI library loads without a GPU and exports exactly what include/leaf_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "leaf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(leaf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from leaf_b200 import _native, build
    build.build()
    L = ctypes.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_native.SIGNATURES) == names
    assert L.leaf_version is not None


def test_no_cpu_path_without_device():
    """On a box without a GPU the engine must refuse loudly instead of computing anything on the host."""
    import torch
    if torch.cuda.is_available():
        return
    from leaf_b200 import _native
    L = _native.lib()
    cfg = _native.LeafCfg(128, 2, 2, 64, 0, 1e-5)
    h = ctypes.c_void_p()
    rc = L.leaf_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -2 and b"no CUDA device" in L.leaf_last_error()
    import pytest
    from leaf_b200 import LeafError, LeafEngine
    with pytest.raises(LeafError):
        LeafEngine({}, heads=2)
