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

    # NSin (D2Q9 only)
    def nsin_init(self, l, rho, ux, uy, uz):
        pl.NSin.InitialCondition(l.p, up(rho), up(ux), up(uy))

    def nsin_macro_collide(self, l, rho, ux, uy, uz, nu, issave):
        host = [rho, ux, uy]
        dev = self._macros(host)
        pl.NSin.MacroCollide(l.p, *dev, nu, bool(issave))
        if issave:
            for h, d in zip(host, dev):
                d.to_host(h)

    def nsin_macro_brinkman_collide(self, l, rho, ux, uy, uz, nu, alpha, issave):
        host = [rho, ux, uy]
        dev = self._macros(host)
        pl.NSin.MacroBrinkmanCollide(l.p, *dev, nu, up(alpha), bool(issave))
        if issave:
            for h, d in zip(host, dev):
                d.to_host(h)

    def nsin_bc_set_u(self, l, uxg, uyg, uzg, mask):
        pl.NSin.BoundaryConditionSetU(l.p, l.dense(uxg), l.dense(uyg), l.dense(mask))

    def nsin_bc_set_rho(self, l, v0, v1, v2, mask):
        pl.NSin.BoundaryConditionSetRho(l.p, l.dense(v0), l.dense(v1), l.dense(mask))

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


# ---- the rest of the op-level interface (AD / ANS / AAD, sensitivities), added as methods below ------------
def _dl(host_arrays, dev_arrays):
    for h, d in zip(host_arrays, dev_arrays):
        if h is not None and d is not None:
            d.to_host(h)


def _vec(self, x, y, z):
    return [x, y, z] if self.dim == 3 else [x, y]


def _snap_dev(l, snap_host):
    return pl.DeviceArray(l.nxyz*l.nc, 0.0) if snap_host is not None else None


def _snap_back(l, dev, host, issave):
    if dev is not None and issave:
        host[:] = api.snapshot_to_host(l.p, dev)


def _snap_up(l, host_ref_layout):
    """reference `_g` layout -> device SoA [c][nxyz]"""
    n, nc = l.nxyz, l.nc
    npk = 4*(n//4)
    soa = np.empty((nc, n))
    if npk:
        soa[:, :npk] = host_ref_layout[:npk*nc].reshape(npk//4, nc, 4).transpose(1, 0, 2).reshape(nc, npk)
    if n > npk:
        soa[:, npk:] = host_ref_layout[npk*nc:].reshape(n - npk, nc).T
    return pl.DeviceArray.from_host(soa)


class _Ext:
    def ad_init(self, l, tem, ux, uy, uz):
        pl.AD.InitialCondition(l.p, up(tem), *[up(a) for a in _vec(self, ux, uy, uz)])

    def ans_init(self, l, ux, uy, uz, ip, iux, iuy, iuz):
        pl.ANS.InitialCondition(l.p, *[up(a) for a in _vec(self, ux, uy, uz)], up(ip), *[up(a) for a in _vec(self, iux, iuy, iuz)])

    def aad_init(self, l, ux, uy, uz, item, iqx, iqy, iqz):
        pl.AAD.InitialCondition(l.p, *[up(a) for a in _vec(self, ux, uy, uz)], up(item), *[up(a) for a in _vec(self, iqx, iqy, iqz)])

    # -- forward two-lattice collides
    def _fwd(self, fn, f, rho, ux, uy, uz, pre, g, tem, qx, qy, qz, post, issave, snap=None, has_snap=False):
        hf = [rho] + _vec(self, ux, uy, uz); hq = [tem] + _vec(self, qx, qy, qz)
        df, dq = [up(a) for a in hf], [up(a) for a in hq]
        sd = _snap_dev(f, snap) if has_snap else None
        extra = [bool(issave)] + ([sd] if has_snap else [])
        fn(f.p, *df, *pre, g.p, *dq, *post, *extra)
        if issave:
            _dl(hf, df); _dl(hq, dq)
        _snap_back(f, sd, snap, issave)

    def ad_macro_collide_force_convection(self, f, rho, ux, uy, uz, nu, g, tem, qx, qy, qz, k, issave):
        self._fwd(pl.AD.MacroCollideForceConvection, f, rho, ux, uy, uz, [nu], g, tem, qx, qy, qz, [k], issave)

    def ad_macro_collide_natural_convection(self, f, rho, ux, uy, uz, nu, g, tem, qx, qy, qz, k, gx, gy, gz, tem0, issave):
        self._fwd(pl.AD.MacroCollideNaturalConvection, f, rho, ux, uy, uz, [nu], g, tem, qx, qy, qz, [k] + _vec(self, gx, gy, gz) + [tem0], issave)

    def ad_macro_brinkman_collide_heat_exchange(self, f, rho, ux, uy, uz, alpha, nu, g, tem, qx, qy, qz, beta, k, issave):
        self._fwd(pl.AD.MacroBrinkmanCollideHeatExchange, f, rho, ux, uy, uz, [up(alpha), nu], g, tem, qx, qy, qz, [up(beta), k], issave)

    def ad_macro_brinkman_collide_force_convection(self, f, rho, ux, uy, uz, alpha, nu, g, tem, qx, qy, qz, kappa, issave, snap):
        self._fwd(pl.AD.MacroBrinkmanCollideForceConvection, f, rho, ux, uy, uz, [up(alpha), nu], g, tem, qx, qy, qz, [up(kappa)], issave, snap, True)

    def ad_macro_brinkman_collide_natural_convection(self, f, rho, ux, uy, uz, alpha, nu, g, tem, qx, qy, qz, kappa, gx, gy, gz, tem0, issave, snap):
        self._fwd(pl.AD.MacroBrinkmanCollideNaturalConvection, f, rho, ux, uy, uz, [up(alpha), nu], g, tem, qx, qy, qz,
                  [up(kappa)] + _vec(self, gx, gy, gz) + [tem0], issave, snap, True)

    # -- adjoint collides
    def ans_macro_brinkman_collide(self, f, rho, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz, nu, alpha, issave):
        ha = [ip] + _vec(self, iux, iuy, iuz) + _vec(self, imx, imy, imz)
        da = [up(a) for a in ha]
        pl.ANS.MacroBrinkmanCollide(f.p, up(rho), *[up(a) for a in _vec(self, ux, uy, uz)], *da, nu, up(alpha), bool(issave))
        if issave:
            _dl(ha, da)

    def _adj(self, fn, f, rho, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz, alpha, nu, g, tem, item, iqx, iqy, iqz, post, issave, snap=None, has_snap=False):
        ha = [ip] + _vec(self, iux, iuy, iuz) + _vec(self, imx, imy, imz); hq = [item] + _vec(self, iqx, iqy, iqz)
        da, dq = [up(a) for a in ha], [up(a) for a in hq]
        sd = _snap_dev(f, snap) if has_snap else None
        extra = [bool(issave)] + ([sd] if has_snap else [])
        fn(f.p, up(rho), *[up(a) for a in _vec(self, ux, uy, uz)], *da, up(alpha), nu, g.p, up(tem), *dq, *post, *extra)
        if issave:
            _dl(ha, da); _dl(hq, dq)
        _snap_back(f, sd, snap, issave)

    def aad_macro_brinkman_collide_heat_exchange(self, f, rho, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz, alpha, nu, g, tem, item, iqx, iqy, iqz, beta, k, issave):
        self._adj(pl.AAD.MacroBrinkmanCollideHeatExchange, f, rho, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz, alpha, nu, g, tem, item, iqx, iqy, iqz,
                  [up(beta), k], issave)

    def aad_macro_brinkman_collide_force_convection(self, f, rho, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz, alpha, nu, g, tem, item, iqx, iqy, iqz, kappa, issave, snap):
        self._adj(pl.AAD.MacroBrinkmanCollideForceConvection, f, rho, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz, alpha, nu, g, tem, item, iqx, iqy, iqz,
                  [up(kappa)], issave, snap, True)

    def aad_macro_brinkman_collide_natural_convection(self, f, rho, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz, alpha, nu, g, tem, item, iqx, iqy, iqz,
                                                      kappa, gx, gy, gz, issave, snap):
        self._adj(pl.AAD.MacroBrinkmanCollideNaturalConvection, f, rho, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz, alpha, nu, g, tem, item, iqx, iqy, iqz,
                  [up(kappa)] + _vec(self, gx, gy, gz), issave, snap, True)

    def aad_macro_brinkman_collide_natural_convection_massflow(self, f, rho, ux, uy, ip, iux, iuy, imx, imy, alpha, nu, g, tem, item, iqx, iqy, kappa,
                                                               gx, gy, dirx, diry, issave, snap):
        self._adj(pl.AAD.MacroBrinkmanCollideNaturalConvectionMassFlow, f, rho, ux, uy, None, ip, iux, iuy, None, imx, imy, None, alpha, nu,
                  g, tem, item, iqx, iqy, None, [up(kappa), gx, gy, up(dirx), up(diry)], issave, snap, True)

    # -- closures
    def ad_bc_set_t(self, g, temg, ux, uy, uz, mask):
        pl.AD.BoundaryConditionSetT(g.p, g.dense(temg), *[up(a) for a in _vec(self, ux, uy, uz)], g.dense(mask))

    def ad_bc_set_q(self, g, qng, ux, uy, uz, kfield, kconst, mask):
        pl.AD.BoundaryConditionSetQ(g.p, g.dense(qng), *[up(a) for a in _vec(self, ux, uy, uz)], up(kfield) if kfield is not None else float(kconst), g.dense(mask))

    def ans_ibc_set_u(self, f, uxg, uyg, uzg, mask, eps):
        pl.ANS.iBoundaryConditionSetU(f.p, *[f.dense(a) for a in _vec(self, uxg, uyg, uzg)], f.dense(mask), eps=eps)

    def ans_ibc_set_rho(self, f, mask):
        pl.ANS.iBoundaryConditionSetRho(f.p, f.dense(mask))

    def aad_ibc_set_t(self, g, ux, uy, uz, mask):
        pl.AAD.iBoundaryConditionSetT(g.p, *[up(a) for a in _vec(self, ux, uy, uz)], g.dense(mask))

    def aad_ibc_set_q(self, g, ux, uy, uz, mask, eps):
        pl.AAD.iBoundaryConditionSetQ(g.p, *[up(a) for a in _vec(self, ux, uy, uz)], g.dense(mask), float(eps))

    def aad_ibc_set_rho(self, f, g, rho, ux, uy, tem, mask, eps):
        pl.AAD.iBoundaryConditionSetRho(f.p, g.p, up(rho), up(ux), up(uy), up(tem), f.dense(mask), eps)

    # -- sensitivities
    def ans_sensitivity_brinkman(self, l, dfds, ux, uy, uz, imx, imy, imz, dads):
        d = up(dfds)
        pl.ANS.SensitivityBrinkman(l.p, d, *[up(a) for a in _vec(self, ux, uy, uz) + _vec(self, imx, imy, imz)], up(dads))
        d.to_host(dfds)

    def aad_sensitivity_heat_exchange(self, l, dfds, ux, uy, uz, imx, imy, imz, dads, tem, item, dbds):
        d = up(dfds)
        pl.AAD.SensitivityHeatExchange(l.p, d, *[up(a) for a in _vec(self, ux, uy, uz) + _vec(self, imx, imy, imz)], up(dads), up(tem), up(item), up(dbds))
        d.to_host(dfds)

    def _bd_args(self, l, ux, uy, uz, imx, imy, imz, dads, tem, item, iqx, iqy, iqz, gs, igs, kappa, dkds):
        return ([up(a) for a in _vec(self, ux, uy, uz) + _vec(self, imx, imy, imz)] + [up(dads), up(tem), up(item)] + [up(a) for a in _vec(self, iqx, iqy, iqz)]
                + [_snap_up(l, gs), _snap_up(l, igs), up(kappa), up(dkds)])

    def aad_sensitivity_brinkman_diffusivity(self, l, dfds, *a):
        d = up(dfds)
        pl.AAD.SensitivityBrinkmanDiffusivity(l.p, d, *self._bd_args(l, *a))
        d.to_host(dfds)

    def aad_sensitivity_temperature_at_heat_source(self, l, dfds, *a):
        *vol, qng, mask = a
        d = up(dfds)
        pl.AAD.SensitivityTemperatureAtHeatSource(l.p, d, *self._bd_args(l, *vol), l.dense(qng), l.dense(mask))
        d.to_host(dfds)


for _k, _v in list(vars(_Ext).items()):
    if not _k.startswith("__"):
        setattr(CudaOps, _k, _v)
