"""CPU-only: the C-ABI library loads and exports every symbol include/panslbm_c.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "panslbm_c.h")
LIB = os.path.join(ROOT, "panslbm2_b200", "libpanslbm_b200.so")


def declared_symbols():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plh?_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        from panslbm2_b200 import build
        build.build()
    return ctypes.CDLL(LIB)


def test_header_declares_a_real_api():
    syms = declared_symbols()
    assert len(syms) >= 40 and "pl_collide" in syms and "pl_plan_advance" in syms


def test_every_declared_symbol_is_exported(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_matches_header():
    from panslbm2_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_no_gpu_fails_loudly(lib):
    """Without a device nothing computes: creating a lattice must fail with a CUDA error, not fall back."""
    lib.pl_device_count.restype = ctypes.c_int
    if lib.pl_device_count() > 0:
        pytest.skip("a GPU is visible here")
    lib.pl_lattice_create.restype = ctypes.c_void_p
    lib.pl_last_error.restype = ctypes.c_char_p
    h = lib.pl_lattice_create(3, 8, 8, 8, 0, 1, 1, 1)
    assert not h
    assert b"cuda" in lib.pl_last_error().lower()
    import panslbm2_b200 as pl
    with pytest.raises(pl.PanslbmError):
        pl.D3Q15(8, 8, 8)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under panslbm2_b200/ may reference it."""
    bad = []
    for d, _, fs in os.walk(os.path.join(ROOT, "panslbm2_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                s = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"lbm_oracle|liblbm_oracle|oracle/_ref|import oracle|from oracle", s):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
