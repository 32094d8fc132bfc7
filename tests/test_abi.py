"""CPU-side checks of the C-ABI boundary: the shared library loads and exports every symbol include/mgp_b200.h
declares (no compute calls -- there is no GPU here), and the product package never imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "manifold_gp_b200", "libmgp_b200.so")
HEADER = os.path.join(ROOT, "include", "mgp_b200.h")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return LIB


def _declared():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mgp_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built):
    dll = ctypes.CDLL(built)
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/mgp_b200.h but not exported by libmgp_b200.so"
    dll.mgp_version.restype = ctypes.c_char_p
    assert b"sm_100a" in dll.mgp_version()


def test_python_binding_covers_the_header(built):
    from manifold_gp_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_argument_validation_without_a_gpu(built):
    """Bad arguments are rejected before any CUDA call, with a message (error behaviour of the boundary)."""
    from manifold_gp_b200 import _lib
    dll = ctypes.CDLL(built)
    dll.mgp_last_error.restype = ctypes.c_char_p
    rc = dll.mgp_lap_spmm_f32(None, None, None, None, None, None, None, None, ctypes.c_int64(4), None, ctypes.c_int64(4),
                              ctypes.c_int64(10), ctypes.c_int32(4), None, None, None, None)
    assert rc == -1 and b"null pointer" in dll.mgp_last_error()
    import torch
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _lib.ptr(torch.zeros(3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "manifold_gp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_no_cpu_fallback_in_operators(built):
    import torch
    import manifold_gp_b200 as mgp
    idx = torch.tensor([[0, 0, 1], [1, 2, 2]])
    val = torch.tensor([0.1, 0.2, 0.3])
    op = mgp.GraphLaplacianOperator(val, idx, 3, torch.tensor([[0.5]]), "symmetric")
    with pytest.raises(RuntimeError, match="CUDA"):
        op._matmul(torch.ones(3, 1))
    with pytest.raises(RuntimeError, match="CUDA"):
        mgp.NearestNeighbors(torch.zeros(10, 3))
