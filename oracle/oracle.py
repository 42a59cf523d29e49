"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes front-end for the two CPU checkers:

* ``Backend("orc")`` — the plain-C restatement ``oracle/liblbm_oracle.so`` (built by
  ``oracle/Makefile`` / ``__graft_entry__.build()``);
* ``Backend("ref", dim)`` — the reference's own OpenMP+AVX headers behind
  ``oracle/_ref/libpanslbm_ref{2d,3d}.so`` (built here from ``/root/reference`` by the same Makefile;
  travels to the GPU box as a prebuilt artefact, never as source).

Both expose the same op-level calls (``ns_macro_collide``, ``stream``, ``bc`` …) so a test can run the same
scenario on either.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may
import this module; the product package ``panslbm2_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORC_PATH = os.path.join(HERE, "liblbm_oracle.so")
REF_PATH = {2: os.path.join(HERE, "_ref", "libpanslbm_ref2d.so"), 3: os.path.join(HERE, "_ref", "libpanslbm_ref3d.so")}
# the same headers compiled WITHOUT _USE_AVX_DEFINES (scalar templates at every site, production/nsopt.cpp:2): Backend("ref_scalar", dim)
REF_SCALAR_PATH = {2: os.path.join(HERE, "_ref", "libpanslbm_ref2d_scalar.so"), 3: os.path.join(HERE, "_ref", "libpanslbm_ref3d_scalar.so")}


def have_ref(dim: int = 3) -> bool:
    return os.path.exists(REF_PATH[dim])


def have_ref_scalar(dim: int = 3) -> bool:
    return os.path.exists(REF_SCALAR_PATH[dim])


def have_orc() -> bool:
    return os.path.exists(ORC_PATH)


def _ptr(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "need C-contiguous ndarray"
    return a.ctypes.data_as(C.c_void_p)


class Lattice:
    def __init__(self, be: "Backend", h, info):
        self.be, self.h = be, h
        (self.lx, self.ly, self.lz, self.peid, self.mx, self.my, self.mz, self.pex, self.pey, self.pez,
         self.nx, self.ny, self.nz, self.nxyz, self.offx, self.offy, self.offz, self.nc) = [int(v) for v in info]

    def get(self):
        f0 = np.empty(self.nxyz)
        f = np.empty(self.nxyz * (self.nc - 1))
        self.be._call("lattice_get", self.h, f0, f)
        return f0, f

    def set(self, f0, f):
        self.be._call("lattice_set", self.h, np.ascontiguousarray(f0, dtype=np.float64), np.ascontiguousarray(f, dtype=np.float64))

    def free(self):
        if self.h is not None:
            self.be._call("lattice_destroy", self.h)
            self.h = None


class Backend:
    """kind='orc', 'ref' or 'ref_scalar'.  dim is 2 or 3 (the reference build has one library per lattice)."""

    def __init__(self, kind: str, dim: int = 3):
        self.kind, self.dim = kind, dim
        path = ORC_PATH if kind == "orc" else (REF_SCALAR_PATH[dim] if kind == "ref_scalar" else REF_PATH[dim])
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.prefix = "orc_" if kind == "orc" else "ref_"

    def has(self, name: str) -> bool:
        return hasattr(self.lib, self.prefix + name)

    def _fn(self, name, restype=None):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        return fn

    def _call(self, name, *args, restype=None):
        conv = []
        for a in args:
            if isinstance(a, np.ndarray) or a is None:
                conv.append(_ptr(a))
            elif isinstance(a, Lattice):
                conv.append(C.c_void_p(a.h))
            elif isinstance(a, float):
                conv.append(C.c_double(a))
            elif isinstance(a, (int, np.integer)) and not isinstance(a, bool):
                conv.append(C.c_int(int(a)))
            elif isinstance(a, bool):
                conv.append(C.c_int(int(a)))
            else:
                conv.append(a)
        return self._fn(name, restype)(*conv)

    def lattice(self, lx, ly, lz=1, peid=0, mx=1, my=1, mz=1) -> Lattice:
        if self.kind == "orc":
            h = self._call("lattice_create", self.dim, lx, ly, lz, peid, mx, my, mz, restype=C.c_void_p)
        else:
            h = self._call("lattice_create", lx, ly, lz, peid, mx, my, mz, restype=C.c_void_p)
        info = np.zeros(18, dtype=np.int32)
        self._fn("lattice_info")(C.c_void_p(h), _ptr(info))
        return Lattice(self, h, info)

    def __getattr__(self, name):
        # generic op call:  be.ns_macro_collide(lat, rho, ux, uy, uz, nu, issave)
        if name.startswith("_"):
            raise AttributeError(name)
        restype = C.c_double if name.startswith(("residual", "time_")) else None

        def call(*args):
            return self._call(name, *args, restype=restype)

        return call
