"""The C-ABI library loads and exports every symbol include/slb200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "slb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(slb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound():
    from slb200 import _lib

    names = _declared()
    assert len(names) >= 30
    assert os.path.exists(_lib.LIB_PATH), "build libslb200.so first (__graft_entry__.build())"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/slb200.h but not exported"
    bound = {s[0] for s in _lib.SIGNATURES}
    assert bound == set(names), (bound ^ set(names))


def test_no_torch_types_in_signatures():
    txt = open(os.path.join(ROOT, "include", "slb200.h")).read()
    assert 'extern "C"' in txt
    code = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code


def test_fails_loudly_without_gpu():
    """No CPU fallback: without a device the context cannot be created."""
    from slb200 import _lib

    if _lib.lib().slb_device_count() > 0:
        pytest.skip("a GPU is present")
    import slb200 as S

    with pytest.raises(S.SlbError):
        S.default_context()
    h = ctypes.c_void_p()
    rc = _lib.lib().slb_ctx_create(0, None, ctypes.byref(h))
    assert rc == -2 and b"no CPU fallback" in _lib.lib().slb_last_error()


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "semilagrangian.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".jl", ".sh")):
                src = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, fn
