"""Op-level parity scenarios shared by every backend: the reference build ("ref", oracle/_ref), the C restatement
("orc"), the product's site math compiled for the host ("hm", tests/hostmath) and the CUDA path ("cuda",
tests/cuda_ops.py).  A scenario takes a backend and returns a list of (name, array) results; two backends agree when
all arrays are equal bit for bit.  Inputs are seeded numpy draws; nothing is read from disk."""
import numpy as np

from helpers import gcoords, i32, random_field, random_pops

NU = 0.07


class Fields:
    """forward / adjoint macroscopic inputs on one block"""

    def __init__(self, n, seed):
        r = lambda k, lo, hi: random_field(n, seed*100 + k, lo, hi)
        self.rho = r(1, 0.9, 1.1); self.ux = r(2, -0.1, 0.1); self.uy = r(3, -0.1, 0.1); self.uz = r(4, -0.1, 0.1)
        self.tem = r(5, 0.0, 1.0)
        self.alpha = r(6, 0.0, 40.0); self.kappa = r(7, 0.02, 0.3); self.beta = r(8, 0.0, 0.5)
        self.dirx = r(9, -1, 1); self.diry = r(10, -1, 1)
        self.dads = r(11, -5, 0); self.dkds = r(12, -1, 1); self.dbds = r(13, -1, 1)
        self.imx = r(14, -0.1, 0.1); self.imy = r(15, -0.1, 0.1); self.imz = r(16, -0.1, 0.1)
        self.item = r(17, -1, 1); self.iqx = r(18, -0.1, 0.1); self.iqy = r(19, -0.1, 0.1); self.iqz = r(20, -0.1, 0.1)


def out(n, k):
    return [np.full(n, -7.0 - i) for i in range(k)]


def two_lattices(be, size, peid, m, seed):
    f = be.lattice(*size, peid, *m)
    g = be.lattice(*size, peid, *m)
    f.set(*random_pops(f.nxyz, f.nc, seed))
    g.set(*random_pops(g.nxyz, g.nc, seed + 1000))
    return f, g


def pops(tag, l):
    a, b = l.get()
    return [(tag + ".f0", a), (tag + ".f", b)]


# ---------------------------------------------------------------------------------------------------------
# collides: every model twice (issave on, then off with another viscosity) from random populations
FORWARD_MODELS = ["ad_force_convection", "ad_natural_convection", "ad_heat_exchange", "ad_brinkman_force_convection", "ad_brinkman_natural_convection"]
ADJOINT_MODELS = ["ans_brinkman", "aad_heat_exchange", "aad_force_convection", "aad_natural_convection", "aad_natural_convection_massflow"]


def collide(be, dim, model, size, seed, peid=0, m=(1, 1, 1)):
    f, g = two_lattices(be, size, peid, m, seed)
    n = f.nxyz
    F = Fields(n, seed)
    G = (0.0, -1.6e-3, 3.0e-4)
    res = []
    snap = np.zeros(n*f.nc)
    for issave, nu in ((1, NU), (0, 0.021)):
        rho, ux, uy, uz, tem, qx, qy, qz = out(n, 8)
        ip, iux, iuy, iuz, imx, imy, imz, item, iqx, iqy, iqz = out(n, 11)
        if model == "ad_force_convection":
            be.ad_macro_collide_force_convection(f, rho, ux, uy, uz, nu, g, tem, qx, qy, qz, 0.11, issave)
        elif model == "ad_natural_convection":
            be.ad_macro_collide_natural_convection(f, rho, ux, uy, uz, nu, g, tem, qx, qy, qz, 0.11, *G, 0.5, issave)
        elif model == "ad_heat_exchange":
            be.ad_macro_brinkman_collide_heat_exchange(f, rho, ux, uy, uz, F.alpha, nu, g, tem, qx, qy, qz, F.beta, 0.11, issave)
        elif model == "ad_brinkman_force_convection":
            be.ad_macro_brinkman_collide_force_convection(f, rho, ux, uy, uz, F.alpha, nu, g, tem, qx, qy, qz, F.kappa, issave, snap)
        elif model == "ad_brinkman_natural_convection":
            be.ad_macro_brinkman_collide_natural_convection(f, rho, ux, uy, uz, F.alpha, nu, g, tem, qx, qy, qz, F.kappa, *G, 0.5, issave, snap)
        elif model == "ans_brinkman":
            be.ans_macro_brinkman_collide(f, F.rho, F.ux, F.uy, F.uz, ip, iux, iuy, iuz, imx, imy, imz, nu, F.alpha, issave)
        elif model == "aad_heat_exchange":
            be.aad_macro_brinkman_collide_heat_exchange(f, F.rho, F.ux, F.uy, F.uz, ip, iux, iuy, iuz, imx, imy, imz, F.alpha, nu,
                                                        g, F.tem, item, iqx, iqy, iqz, F.beta, 0.11, issave)
        elif model == "aad_force_convection":
            be.aad_macro_brinkman_collide_force_convection(f, F.rho, F.ux, F.uy, F.uz, ip, iux, iuy, iuz, imx, imy, imz, F.alpha, nu,
                                                           g, F.tem, item, iqx, iqy, iqz, F.kappa, issave, snap)
        elif model == "aad_natural_convection":
            be.aad_macro_brinkman_collide_natural_convection(f, F.rho, F.ux, F.uy, F.uz, ip, iux, iuy, iuz, imx, imy, imz, F.alpha, nu,
                                                             g, F.tem, item, iqx, iqy, iqz, F.kappa, *G, issave, snap)
        elif model == "aad_natural_convection_massflow":
            assert dim == 2
            be.aad_macro_brinkman_collide_natural_convection_massflow(f, F.rho, F.ux, F.uy, ip, iux, iuy, imx, imy, F.alpha, nu,
                                                                      g, F.tem, item, iqx, iqy, F.kappa, G[0], G[1], F.dirx, F.diry, issave, snap)
        else:
            raise KeyError(model)
        if issave:
            if model.startswith("ad_"):
                arrs = [rho, ux, uy, uz, tem, qx, qy, qz]
                names = ["rho", "ux", "uy", "uz", "tem", "qx", "qy", "qz"]
            elif model == "ans_brinkman":
                arrs = [ip, iux, iuy, iuz, imx, imy, imz]
                names = ["ip", "iux", "iuy", "iuz", "imx", "imy", "imz"]
            else:
                arrs = [ip, iux, iuy, iuz, imx, imy, imz, item, iqx, iqy, iqz]
                names = ["ip", "iux", "iuy", "iuz", "imx", "imy", "imz", "item", "iqx", "iqy", "iqz"]
            for nm, a in zip(names, arrs):
                if dim == 2 and nm.endswith("z"):
                    continue
                res.append((nm, a.copy()))
            if model in ("ad_brinkman_force_convection", "ad_brinkman_natural_convection", "aad_force_convection", "aad_natural_convection",
                         "aad_natural_convection_massflow"):
                res.append(("snapshot", snap.copy()))
        res += pops(f"f{issave}", f)
        if model != "ans_brinkman":
            res += pops(f"g{issave}", g)
    f.free(); g.free()
    return res


# ---------------------------------------------------------------------------------------------------------
# closures on the global boundary planes, random masks and values
CLOSURES = ["ad_set_t", "ad_set_q_const", "ad_set_q_field", "ans_iset_u", "ans_iset_rho", "aad_iset_t", "aad_iset_q", "aad_iset_rho"]


def closure(be, dim, kind, size, seed, peid=0, m=(1, 1, 1)):
    f, g = two_lattices(be, size, peid, m, seed)
    n = f.nxyz
    F = Fields(n, seed)
    Gn = size[0]*size[1]*size[2]
    rs = np.random.RandomState(seed + 7)
    mask = i32(rs.randint(0, 2, size=Gn))
    mask3 = i32(rs.randint(0, 3, size=Gn))
    v = [random_field(Gn, seed*10 + d, -0.1, 0.1) for d in range(3)]
    tg = random_field(Gn, seed*10 + 4, 0.0, 1.0)
    if kind == "ad_set_t":
        be.ad_bc_set_t(g, tg, F.ux, F.uy, F.uz, mask)
    elif kind == "ad_set_q_const":
        be.ad_bc_set_q(g, tg, F.ux, F.uy, F.uz, None, 0.13, mask)
    elif kind == "ad_set_q_field":
        be.ad_bc_set_q(g, tg, F.ux, F.uy, F.uz, F.kappa, 0.0, mask)
    elif kind == "ans_iset_u":
        be.ans_ibc_set_u(f, v[0], v[1], v[2], mask, 0.0)
        be.ans_ibc_set_u(f, v[1], v[2], v[0], mask, 1.0)
    elif kind == "ans_iset_rho":
        be.ans_ibc_set_rho(f, mask)
    elif kind == "aad_iset_t":
        be.aad_ibc_set_t(g, F.ux, F.uy, F.uz, mask)
    elif kind == "aad_iset_q":
        be.aad_ibc_set_q(g, F.ux, F.uy, F.uz, mask, 0.0)
        be.aad_ibc_set_q(g, F.uy, F.uz, F.ux, mask, 1.0)
    elif kind == "aad_iset_rho":
        assert dim == 2
        be.aad_ibc_set_rho(f, g, F.rho, F.ux, F.uy, F.tem, mask3, 0.0)
        be.aad_ibc_set_rho(f, g, F.rho, F.uy, F.ux, F.tem, mask3, 1.0)
    else:
        raise KeyError(kind)
    res = pops("f", f) + pops("g", g)
    f.free(); g.free()
    return res


# ---------------------------------------------------------------------------------------------------------
SENSITIVITIES = ["ans_brinkman", "aad_heat_exchange", "aad_brinkman_diffusivity", "aad_temperature_at_heat_source"]


def ref_snapshot(n, nc, seed):
    """a random population snapshot in the reference's `_g` layout; every backend converts from this"""
    return random_field(n*nc, seed, 0.0, 0.2)


def sensitivity(be, dim, kind, size, seed, peid=0, m=(1, 1, 1)):
    g = be.lattice(*size, peid, *m)
    n = g.nxyz
    F = Fields(n, seed)
    Gn = size[0]*size[1]*size[2]
    dfds = random_field(n, seed + 31, -1, 1)
    gs, igs = ref_snapshot(n, g.nc, seed + 32), ref_snapshot(n, g.nc, seed + 33)
    mask = i32(np.random.RandomState(seed + 34).randint(0, 2, size=Gn))
    qn = random_field(Gn, seed + 35, 0.0, 0.02)
    if kind == "ans_brinkman":
        be.ans_sensitivity_brinkman(g, dfds, F.ux, F.uy, F.uz, F.imx, F.imy, F.imz, F.dads)
    elif kind == "aad_heat_exchange":
        be.aad_sensitivity_heat_exchange(g, dfds, F.ux, F.uy, F.uz, F.imx, F.imy, F.imz, F.dads, F.tem, F.item, F.dbds)
    elif kind == "aad_brinkman_diffusivity":
        be.aad_sensitivity_brinkman_diffusivity(g, dfds, F.ux, F.uy, F.uz, F.imx, F.imy, F.imz, F.dads, F.tem, F.item, F.iqx, F.iqy, F.iqz,
                                                gs, igs, F.kappa, F.dkds)
    elif kind == "aad_temperature_at_heat_source":
        be.aad_sensitivity_temperature_at_heat_source(g, dfds, F.ux, F.uy, F.uz, F.imx, F.imy, F.imz, F.dads, F.tem, F.item, F.iqx, F.iqy, F.iqz,
                                                      gs, igs, F.kappa, F.dkds, qn, mask)
    else:
        raise KeyError(kind)
    g.free()
    return [("dfds", dfds)]


def inits(be, dim, size, seed, peid=0, m=(1, 1, 1)):
    res = []
    for fam in ("ns", "ad", "ans", "aad"):
        l = be.lattice(*size, peid, *m)
        F = Fields(l.nxyz, seed)
        if fam == "ns":
            be.ns_init(l, F.rho, F.ux, F.uy, F.uz)
        elif fam == "ad":
            be.ad_init(l, F.tem, F.ux, F.uy, F.uz)
        elif fam == "ans":
            be.ans_init(l, F.ux, F.uy, F.uz, F.item, F.imx, F.imy, F.imz)
        else:
            be.aad_init(l, F.ux, F.uy, F.uz, F.item, F.iqx, F.iqy, F.iqz)
        res += pops(fam, l)
        l.free()
    return res


# ---------------------------------------------------------------------------------------------------------
# NSin (src/equation/nsincompressible.h; D2Q9 only): InitialCondition, both collides with and without issave, SetU and SetRho on all
# four edges with random masks and values, and a short loop collide - Stream - bounce - SetU - SetRho - SmoothCorner
def nsin(be, size, seed, steps=6):
    lx, ly = size[0], size[1]
    l = be.lattice(lx, ly, 1)
    n = l.nxyz
    F = Fields(n, seed)
    res = []
    be.nsin_init(l, F.rho, F.ux, F.uy, F.uz)
    res += pops("init", l)
    l.set(*random_pops(n, l.nc, seed))
    for k, (issave, nu, alpha) in enumerate(((1, NU, None), (0, 0.021, None), (1, 0.1, F.alpha), (0, 0.03, F.alpha))):
        rho, ux, uy, uz = out(n, 4)
        if alpha is None:
            be.nsin_macro_collide(l, rho, ux, uy, uz, nu, issave)
        else:
            be.nsin_macro_brinkman_collide(l, rho, ux, uy, uz, nu, alpha, issave)
        if issave:
            res += [(f"rho{k}", rho.copy()), (f"ux{k}", ux.copy()), (f"uy{k}", uy.copy())]
        res += pops(f"c{k}", l)
    Gn = lx*ly
    rs = np.random.RandomState(seed + 7)
    mask = i32(rs.randint(0, 2, size=Gn))
    v = [random_field(Gn, seed*10 + d, -0.1, 0.1) for d in range(2)]
    rg = random_field(Gn, seed*10 + 3, 0.95, 1.05)
    be.nsin_bc_set_u(l, v[0], v[1], None, mask)
    res += pops("setu", l)
    be.nsin_bc_set_rho(l, rg, v[1], None, mask)
    res += pops("setrho", l)
    if steps <= 0:      # backends without Stream / SmoothCorner (tests/hostmath: site math only)
        l.free()
        return res
    # a short driver-style loop: lid-driven box with a pressure outlet patch on ymin
    i, j, _ = gcoords(lx, ly, 1)
    wall = i32(np.where((i == 0) | (i == lx - 1) | ((j == 0) & (i < lx//2)), 1, 0))
    lid = i32(j == ly - 1)
    outlet = i32((j == 0) & (i >= lx//2))
    uxg, uyg = np.full(Gn, 0.05), np.zeros(Gn)
    rho, ux, uy, uz = np.ones(n), np.zeros(n), np.zeros(n), np.zeros(n)
    be.nsin_init(l, rho, ux, uy, uz)
    for _ in range(steps):
        be.nsin_macro_brinkman_collide(l, rho, ux, uy, uz, 0.1, F.alpha*0.01, 1)
        be.stream(l)
        be.bc(l, wall, 0)
        be.nsin_bc_set_u(l, uxg, uyg, None, lid)
        be.nsin_bc_set_rho(l, np.ones(Gn), uyg, None, outlet)
        be.smooth_corner(l)
    res += [("loop.rho", rho.copy()), ("loop.ux", ux.copy()), ("loop.uy", uy.copy())] + pops("loop", l)
    l.free()
    return res


def assert_same(ra, rb, what=""):
    assert len(ra) == len(rb)
    for (na, a), (nb, b) in zip(ra, rb):
        assert na == nb
        assert np.array_equal(a, b), f"{what}: {na} differs (max abs diff {np.max(np.abs(a - b)):.3e})"
