"""The C-ABI shared library loads and exports every symbol include/detex_b200.h declares.
No compute calls: on a box without a GPU dtx_create must fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from detex_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "detex_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dtx_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    for name in declared_symbols():
        assert hasattr(L, name), name
    assert L.dtx_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.load()
    h = ctypes.c_void_p()
    rc = L.dtx_create(0, None, ctypes.byref(h))
    assert rc != 0 and not h.value
    from detex_b200.engine import DtxError, Engine
    with pytest.raises(DtxError):
        Engine(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under detex_b200/ may import it."""
    pkg = os.path.join(ROOT, "detex_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "detex_oracle" not in txt and "ref_shim" not in txt, f
