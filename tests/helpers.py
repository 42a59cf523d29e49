"""Shared deterministic scenario builders for the parity tests (closed-form, seedless where the reference's
own fixtures are; numpy RandomState with fixed seeds where a stress state is wanted)."""
import numpy as np

W3 = np.array([2/9] + [1/9]*6 + [1/72]*8)
W2 = np.array([4/9] + [1/9]*4 + [1/36]*4)


def random_pops(nxyz, nc, seed, amp=0.05):
    """Positive populations near the rest equilibrium, in the reference layout (f0[nxyz], f[(nc-1)*idx+c-1])."""
    rs = np.random.RandomState(seed)
    w = W3 if nc == 15 else W2
    p = w[None, :] * (1.0 + amp * rs.uniform(-1, 1, size=(nxyz, nc)))
    return np.ascontiguousarray(p[:, 0]), np.ascontiguousarray(p[:, 1:].reshape(-1))


def random_field(n, seed, lo, hi):
    return np.random.RandomState(seed).uniform(lo, hi, size=n)


def gcoords(lx, ly, lz=1):
    """global coordinate arrays flattened in g = i + lx*(j + ly*k) order"""
    k, j, i = np.meshgrid(np.arange(lz), np.arange(ly), np.arange(lx), indexing="ij")
    return i.reshape(-1), j.reshape(-1), k.reshape(-1)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def same(a, b):
    """bit-exact up to the sign of zero"""
    return np.array_equal(np.asarray(a), np.asarray(b))


def relinf(a, b):
    a, b = np.asarray(a), np.asarray(b)
    d = np.max(np.abs(a - b)) if a.size else 0.0
    s = np.max(np.abs(b)) if b.size else 0.0
    return d / s if s > 0 else d


def hostmath_backend(dim):
    """the product's site math compiled for the host (tests/hostmath), driven like an oracle backend"""
    import ctypes as C
    import importlib.util
    import os
    from oracle import oracle as O
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("hostmath_build", os.path.join(here, "hostmath", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class HostMath(O.Backend):
        def __init__(self, dim):
            self.kind, self.dim = "hm", dim
            self.lib = C.CDLL(mod.build())
            self.prefix = "hm_"

        def lattice(self, lx, ly, lz=1, peid=0, mx=1, my=1, mz=1):
            h = self._call("lattice_create", self.dim, lx, ly, lz, peid, mx, my, mz, restype=C.c_void_p)
            info = np.zeros(18, dtype=np.int32)
            self._fn("lattice_info")(C.c_void_p(h), info.ctypes.data_as(C.c_void_p))
            return O.Lattice(self, h, info)

    return HostMath(dim)
