"""CPU: the C-ABI library loads and exports exactly the symbols include/rdm_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "rdm_b200.h")).read()
    return sorted(set(re.findall(r"RDM_API[^;(]*?\b(rdm_\w+)\s*\(", src)))


def test_header_declares_symbols():
    syms = _header_symbols()
    assert "rdm_knn_search" in syms and "rdm_last_error" in syms and len(syms) >= 10


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from rdm_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(L, name), f"{name} declared in include/rdm_b200.h but not exported"
    assert L.rdm_abi_version() >= 1


def test_python_binding_table_matches_header():
    from rdm_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_every_declaration_cites_the_reference():
    src = open(os.path.join(ROOT, "include", "rdm_b200.h")).read()
    assert len(re.findall(r"\w+\.py:\d+", src)) >= 8


def test_errors_are_reported_not_swallowed():
    from rdm_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    rc = L.rdm_knn_create(ctypes.byref(h), None, 10, 512, 0, 1, 0, 0)     # null db -> argument error, no CUDA call
    assert rc != 0 and b"null" in L.rdm_last_error()
    rc = L.rdm_knn_search(None, None, 1, 4, None, None, None, None)
    assert rc != 0
