"""Adapter that drives the CUDA path (panslbm2_b200.api over the C-ABI) with the same op-level calls and the same
dense-global-array conventions as oracle.oracle.Backend, so one scenario can run on reference / oracle / GPU."""
import numpy as np

import panslbm2_b200 as pl
from panslbm2_b200 import api


class CudaLattice:
    def __init__(self, p):
        self.p = p
        for k in ("lx", "ly", "lz", "nx", "ny", "nz", "nxyz", "nc", "nd"):
            setattr(self, k, getattr(p, k))
        self.offx, self.offy, self.offz = p.offsetx, p.offsety, p.offsetz

    def get(self):
        return self.p.get_populations()

    def set(self, f0, f):
        self.p.set_populations(f0, f)

    def free(self):
        self.p.free()

    def dense(self, arr):
        lx, ly = self.lx, self.ly
        if self.nd == 2:
            return lambda i, j: arr[i + lx*j]
        return lambda i, j, k: arr[i + lx*(j + ly*k)]


def up(a):
    return None if a is None else pl.DeviceArray.from_host(a)


class CudaOps:
    kind = "cuda"

    def __init__(self, dim=3):
        self.dim = dim

    def has(self, name):
        return hasattr(self, name)

    def lattice(self, lx, ly, lz=1, peid=0, mx=1, my=1, mz=1):
        p = pl.D2Q9(lx, ly, peid, mx, my) if self.dim == 2 else pl.D3Q15(lx, ly, lz, peid, mx, my, mz)
        return CudaLattice(p)

    # particle ops
    def stream(self, l): l.p.Stream()
    def istream(self, l): l.p.iStream()
    def smooth_corner(self, l): l.p.SmoothCorner()

    def bc(self, l, bct, inverse):
        (l.p.iBoundaryCondition if inverse else l.p.BoundaryCondition)(l.dense(bct))

    def bc_plane(self, l, axis, coord, d, bct, inverse):
        l.p._bounce_plane(axis, coord, d, l.dense(bct), bool(inverse))

    # NS
    def _z(self, a):
        return a if self.dim == 3 else None

    def ns_init(self, l, rho, ux, uy, uz):
        pl.NS.InitialCondition(l.p, up(rho), up(ux), up(uy), up(self._z(uz)))

    def _macros(self, arrs):
        return [up(a) if a is not None else None for a in arrs]

    def ns_macro_collide(self, l, rho, ux, uy, uz, nu, issave):
        host = [rho, ux, uy] + ([uz] if self.dim == 3 else [])
        dev = self._macros(host)
        pl.NS.MacroCollide(l.p, *dev, nu, bool(issave))
        if issave:
            for h, d in zip(host, dev):
                d.to_host(h)

    def ns_macro_brinkman_collide(self, l, rho, ux, uy, uz, nu, alpha, issave):
        host = [rho, ux, uy] + ([uz] if self.dim == 3 else [])
        dev = self._macros(host)
        pl.NS.MacroBrinkmanCollide(l.p, *dev, nu, up(alpha), bool(issave))
        if issave:
            for h, d in zip(host, dev):
                d.to_host(h)

    def ns_bc_set_u(self, l, uxg, uyg, uzg, mask):
        fns = [l.dense(uxg), l.dense(uyg)] + ([l.dense(uzg)] if self.dim == 3 else [])
        pl.NS.BoundaryConditionSetU(l.p, *fns, l.dense(mask))

    def ns_bc_set_rho(self, l, v0, v1, v2, mask):
        fns = [l.dense(v0), l.dense(v1)] + ([l.dense(v2)] if self.dim == 3 else [])
        pl.NS.BoundaryConditionSetRho(l.p, *fns, l.dense(mask))

    # utilities
    def residual3(self, ux, uy, uz, uxp, uyp, uzp, n):
        return pl.Residual(up(ux), up(uy), up(uz), up(uxp), up(uyp), up(uzp), n)

    def residual2(self, ux, uy, uxp, uyp, n):
        return pl.Residual(up(ux), up(uy), up(uxp), up(uyp), n)

    def residual1(self, ux, uxp, n):
        return pl.Residual(up(ux), up(uxp), n)

    def normalize(self, v, n):
        d = up(v)
        pl.Normalize(d, n)
        d.to_host(v)
