"""One optimisation iteration of production/heatsink.cpp (D2Q9) / production/heatsink3D.cpp (D3Q15) without the filter and
MMA stages: design -> alpha, diffusivity -> forward loop (heatsink3D.cpp:148-184) -> adjoint loop (:191-224) ->
sensitivity (:241-246), with a fixed step budget (the convergence `break` disabled as in test/nsadncsens.cpp:93-96) so that
every backend executes the same number of steps.  Written twice over the same parameters: `run_oplevel` drives an
oracle-style backend (reference build / C restatement), `run_cuda` drives the product through the Python mirror of the
reference API, call by call or through the fused plan."""
import math

import numpy as np

from helpers import gcoords, i32

FWD = ["rho", "ux", "uy", "uz", "tem", "qx", "qy", "qz"]
ADJ = ["ip", "iux", "iuy", "iuz", "imx", "imy", "imz", "item", "iqx", "iqy", "iqz"]


def params(dim, size):
    lx, ly, lz = size
    p = dict(dim=dim, lx=lx, ly=ly, lz=lz, Pr=6.0, Ra=1e3 if dim == 3 else 1e4, nu=0.1, tem0=0.0, qn0=1.5e-2 if dim == 3 else 1.0e-2,
             alphamax=1.0e4, qg=1.0, qf=1e-2)
    p["L"] = max(2.0, float((lx - 1)//4))
    p["mx"], p["my"], p["mz"] = 3*(lx - 1)//4 + 1, 3*(ly - 1)//4 + 1, (3*(lz - 1)//4 + 1) if dim == 3 else 1
    U = p["nu"]*math.sqrt(p["Ra"]/p["Pr"])/float(ly - 1)
    p["diff_fluid"] = p["nu"]/p["Pr"]
    p["diff_solid"] = p["diff_fluid"]*10.0
    p["gx"], p["gy"], p["gz"] = 0.0, U*U/float(ly - 1), 0.0
    return p


def design_variable(p, i, j, k):
    """closed-form grey (filtered) design on global coordinates: what the filter stage hands to heatsink3D.cpp:114-119"""
    inbox = (i < p["mx"]) & (j < p["my"]) & ((k < p["mz"]) if p["dim"] == 3 else True)
    return np.where(inbox, 0.5 + 0.4*np.sin(0.37*i)*np.cos(0.23*j)*np.sin(0.31*k + 0.5), 1.0)


def design_fields(p, i, j, k):
    """closed-form grey design on global coordinates -> alpha, diffusivity, dads, dkds (heatsink3D.cpp:114-119)"""
    ss = design_variable(p, i, j, k)
    qg, qf, ly = p["qg"], p["qf"], p["ly"]
    kappa = p["diff_solid"] + (p["diff_fluid"] - p["diff_solid"])*ss*(1.0 + qg)/(ss + qg)
    alpha = p["alphamax"]/float(ly - 1)*qf*(1.0 - ss)/(ss + qf)
    dkds = (p["diff_fluid"] - p["diff_solid"])*qg*(1.0 + qg)/(ss + qg)**2
    dads = -p["alphamax"]/float(ly - 1)*qf*(1.0 + qf)/(ss + qf)**2
    return alpha, kappa, dads, dkds


def local_coords(l):
    k, j, i = np.meshgrid(np.arange(l.nz), np.arange(l.ny), np.arange(l.nx), indexing="ij")
    return i.reshape(-1) + l.offx, j.reshape(-1) + l.offy, k.reshape(-1) + l.offz


def predicates(p):
    """the drivers' lambdas as vectorised functions of global coordinates (3-D signature; k ignored in 2-D)"""
    lx, ly, lz, L, d3 = p["lx"], p["ly"], p["lz"], p["L"], p["dim"] == 3
    zmin = (lambda k: k == 0) if d3 else (lambda k: False)
    zmax = (lambda k: k == lz - 1) if d3 else (lambda k: False)
    inL = (lambda i, k: (i < L) & (k < L)) if d3 else (lambda i, k: i < L)
    return dict(
        f_wall=lambda i, j, k: np.where((i == 0) | zmin(k), 2, 1),
        g_wall=lambda i, j, k: np.where((i == 0) | zmin(k), 2, 0),
        setT=lambda i, j, k: (i == lx - 1) | (j == ly - 1) | zmax(k),
        setQ=lambda i, j, k: j == 0,
        source=lambda i, j, k: (j == 0) & inL(i, k),
        qn=lambda i, j, k: np.where((j == 0) & inL(i, k), p["qn0"], 0.0),
        tem=lambda i, j, k: np.full(np.shape(i), p["tem0"]),
    )


def run_oplevel(be, dim, size, nt, peid=0, m=(1, 1, 1), only_forward=False):
    p = params(dim, size)
    f = be.lattice(*size, peid, *m)
    g = be.lattice(*size, peid, *m)
    n = f.nxyz
    alpha, kappa, dads, dkds = [np.ascontiguousarray(a) for a in design_fields(p, *local_coords(f))]
    P = predicates(p)
    gi_, gj_, gk_ = gcoords(*size)
    D = {k: (i32(v(gi_, gj_, gk_)) if k in ("f_wall", "g_wall", "setT", "setQ", "source") else np.ascontiguousarray(v(gi_, gj_, gk_), dtype=np.float64))
         for k, v in P.items()}
    z = lambda: np.zeros(n)
    A = {k: z() for k in FWD + ADJ + ["uxp", "uyp", "uzp", "qxp", "qyp", "qzp", "iuxp", "iuyp", "iuzp", "iqxp", "iqyp", "iqzp"]}
    A["rho"][:] = 1.0
    gsnap, igsnap = np.zeros(n*f.nc), np.zeros(n*f.nc)
    G = (p["gx"], p["gy"], p["gz"])
    be.ns_init(f, A["rho"], A["ux"], A["uy"], A["uz"])
    be.ad_init(g, A["tem"], A["ux"], A["uy"], A["uz"])
    for _ in range(nt):
        be.ad_macro_brinkman_collide_natural_convection(f, A["rho"], A["ux"], A["uy"], A["uz"], alpha, p["nu"], g, A["tem"], A["qx"], A["qy"], A["qz"],
                                                        kappa, *G, p["tem0"], 1, gsnap)
        be.stream(f); be.stream(g)
        be.bc(f, D["f_wall"], 0)
        be.ad_bc_set_t(g, D["tem"], A["ux"], A["uy"], A["uz"], D["setT"])
        be.ad_bc_set_q(g, D["qn"], A["ux"], A["uy"], A["uz"], kappa, 0.0, D["setQ"])
        be.bc(g, D["g_wall"], 0)
        be.smooth_corner(f); be.smooth_corner(g)
        for a, b in (("ux", "uxp"), ("uy", "uyp"), ("uz", "uzp"), ("qx", "qxp"), ("qy", "qyp"), ("qz", "qzp")):
            A[a], A[b] = A[b], A[a]
    res = {k: A[k].copy() for k in FWD}
    res["gsnap"] = gsnap.copy()
    if not only_forward:
        be.ans_init(f, A["ux"], A["uy"], A["uz"], A["ip"], A["iux"], A["iuy"], A["iuz"])
        be.aad_init(g, A["ux"], A["uy"], A["uz"], A["item"], A["iqx"], A["iqy"], A["iqz"])
        for _ in range(nt):
            be.aad_macro_brinkman_collide_natural_convection(f, A["rho"], A["ux"], A["uy"], A["uz"], A["ip"], A["iux"], A["iuy"], A["iuz"],
                                                             A["imx"], A["imy"], A["imz"], alpha, p["nu"], g, A["tem"], A["item"], A["iqx"], A["iqy"], A["iqz"],
                                                             kappa, *G, 1, igsnap)
            be.istream(f); be.istream(g)
            be.aad_ibc_set_t(g, A["ux"], A["uy"], A["uz"], D["setT"])
            be.aad_ibc_set_q(g, A["ux"], A["uy"], A["uz"], D["setQ"], 0.0)
            be.aad_ibc_set_q(g, A["ux"], A["uy"], A["uz"], D["source"], 1.0)
            be.bc(g, D["g_wall"], 1)
            be.bc(f, D["f_wall"], 1)
            be.smooth_corner(f); be.smooth_corner(g)
            for a, b in (("iux", "iuxp"), ("iuy", "iuyp"), ("iuz", "iuzp"), ("iqx", "iqxp"), ("iqy", "iqyp"), ("iqz", "iqzp")):
                A[a], A[b] = A[b], A[a]
        dfdss = np.zeros(n)
        be.aad_sensitivity_temperature_at_heat_source(g, dfdss, A["ux"], A["uy"], A["uz"], A["imx"], A["imy"], A["imz"], dads, A["tem"], A["item"],
                                                      A["iqx"], A["iqy"], A["iqz"], gsnap, igsnap, kappa, dkds, D["qn"], D["source"])
        res.update({k: A[k].copy() for k in ADJ})
        res["igsnap"] = igsnap.copy()
        res["dfdss"] = dfdss
    res["f.f0"], res["f.f"] = f.get()
    res["g.f0"], res["g.f"] = g.get()
    if dim == 2:
        for k in [k for k in res if k.endswith("z")]:
            del res[k]
    f.free(); g.free()
    return res


def time_oplevel(be, size, steps, warmup):
    """seconds of `steps` forward and `steps` adjoint time-loop iterations (after `warmup` each) of the 3-D heatsink loops on a CPU
    backend, op by op as run_oplevel does — bench.py's CPU baseline when only the C restatement is available (kind "port")"""
    import time
    dim = 3
    p = params(dim, size)
    f, g = be.lattice(*size), be.lattice(*size)
    n = f.nxyz
    alpha, kappa, _, _ = [np.ascontiguousarray(a) for a in design_fields(p, *local_coords(f))]
    P = predicates(p)
    gc = gcoords(*size)
    D = {k: (i32(v(*gc)) if k in ("f_wall", "g_wall", "setT", "setQ", "source") else np.ascontiguousarray(v(*gc), dtype=np.float64)) for k, v in P.items()}
    A = {k: np.zeros(n) for k in FWD + ADJ}
    A["rho"][:] = 1.0
    gsnap, igsnap = np.zeros(n*f.nc), np.zeros(n*f.nc)
    G = (p["gx"], p["gy"], p["gz"])
    be.ns_init(f, A["rho"], A["ux"], A["uy"], A["uz"])
    be.ad_init(g, A["tem"], A["ux"], A["uy"], A["uz"])
    secs = [0.0, 0.0]
    for t in range(warmup + steps):
        if t == warmup:
            t0 = time.perf_counter()
        be.ad_macro_brinkman_collide_natural_convection(f, A["rho"], A["ux"], A["uy"], A["uz"], alpha, p["nu"], g, A["tem"], A["qx"], A["qy"], A["qz"],
                                                        kappa, *G, p["tem0"], 1, gsnap)
        be.stream(f); be.stream(g)
        be.bc(f, D["f_wall"], 0)
        be.ad_bc_set_t(g, D["tem"], A["ux"], A["uy"], A["uz"], D["setT"])
        be.ad_bc_set_q(g, D["qn"], A["ux"], A["uy"], A["uz"], kappa, 0.0, D["setQ"])
        be.bc(g, D["g_wall"], 0)
        be.smooth_corner(f); be.smooth_corner(g)
    secs[0] = time.perf_counter() - t0
    be.ans_init(f, A["ux"], A["uy"], A["uz"], A["ip"], A["iux"], A["iuy"], A["iuz"])
    be.aad_init(g, A["ux"], A["uy"], A["uz"], A["item"], A["iqx"], A["iqy"], A["iqz"])
    for t in range(warmup + steps):
        if t == warmup:
            t0 = time.perf_counter()
        be.aad_macro_brinkman_collide_natural_convection(f, A["rho"], A["ux"], A["uy"], A["uz"], A["ip"], A["iux"], A["iuy"], A["iuz"],
                                                         A["imx"], A["imy"], A["imz"], alpha, p["nu"], g, A["tem"], A["item"], A["iqx"], A["iqy"], A["iqz"],
                                                         kappa, *G, 1, igsnap)
        be.istream(f); be.istream(g)
        be.aad_ibc_set_t(g, A["ux"], A["uy"], A["uz"], D["setT"])
        be.aad_ibc_set_q(g, A["ux"], A["uy"], A["uz"], D["setQ"], 0.0)
        be.aad_ibc_set_q(g, A["ux"], A["uy"], A["uz"], D["source"], 1.0)
        be.bc(g, D["g_wall"], 1)
        be.bc(f, D["f_wall"], 1)
        be.smooth_corner(f); be.smooth_corner(g)
    secs[1] = time.perf_counter() - t0
    f.free(); g.free()
    return secs


def run_cuda(dim, size, nt, fused, peid=0, m=(1, 1, 1), only_forward=False, chunks=(1, 3), save_last=None, observe=None):
    """the same iteration through panslbm2_b200's Python mirror of the reference API.  fused=False issues the calls one
    by one exactly like the driver; fused=True records the loop bodies into step plans and advances them in chunks
    (as a driver checking Residual every `dt` steps would).  save_last: only the last k collides of every chunk store their
    macros / snapshot at every site (pl_plan_advance_observed); observe(A, gsnap, igsnap) is called after every chunk, where
    the driver would look at its arrays (heatsink3D.cpp:152-160)."""
    import panslbm2_b200 as pl
    from panslbm2_b200 import api
    p = params(dim, size)
    d3 = dim == 3
    f = pl.D3Q15(*size, peid, *m) if d3 else pl.D2Q9(size[0], size[1], peid, m[0], m[1])
    g = pl.D3Q15(*size, peid, *m) if d3 else pl.D2Q9(size[0], size[1], peid, m[0], m[1])
    n = f.nxyz

    class _L:   # local_coords() wants the oracle lattice attribute names
        nx, ny, nz, offx, offy, offz = f.nx, f.ny, f.nz, f.offsetx, f.offsety, f.offsetz
    alpha, kappa, dads, dkds = [pl.DeviceArray.from_host(a) for a in design_fields(p, *local_coords(_L))]
    P3 = predicates(p)
    P = P3 if d3 else {k: (lambda fn: (lambda i, j: fn(i, j, 0)))(v) for k, v in P3.items()}
    A = {k: pl.DeviceArray(n, 0.0) for k in FWD + ADJ + ["uxp", "uyp", "uzp", "qxp", "qyp", "qzp", "iuxp", "iuyp", "iuzp", "iqxp", "iqyp", "iqzp"]}
    A["rho"].fill(1.0)
    gsnap, igsnap = pl.DeviceArray(n*f.nc, 0.0), pl.DeviceArray(n*f.nc, 0.0)
    V = lambda *names: [A[k] for k in names if d3 or not k.rstrip("p").endswith("z")]
    G = [p["gx"], p["gy"]] + ([p["gz"]] if d3 else [])
    pl.NS.InitialCondition(f, A["rho"], *V("ux", "uy", "uz"))
    pl.AD.InitialCondition(g, A["tem"], *V("ux", "uy", "uz"))

    def swap(pairs):
        for a, b in pairs:
            A[a], A[b] = A[b], A[a]
    fpairs = [("ux", "uxp"), ("uy", "uyp"), ("uz", "uzp"), ("qx", "qxp"), ("qy", "qyp"), ("qz", "qzp")]
    apairs = [("iux", "iuxp"), ("iuy", "iuyp"), ("iuz", "iuzp"), ("iqx", "iqxp"), ("iqy", "iqyp"), ("iqz", "iqzp")]

    def advance(plan, pairs):
        done = 0
        sizes = list(chunks)
        while done < nt:
            c = min(sizes[len(sizes) - 1] if done else sizes[0], nt - done)
            last = done + c == nt
            par0 = plan.parity
            plan.advance(c, end_streamed=last, save_last=save_last)
            done += c
            if observe is not None:
                observe(A, gsnap, igsnap)
        # the driver's pointer state after nt full iterations = nt swaps
        if nt % 2:
            swap(pairs)

    if not fused:
        for _ in range(nt):
            pl.AD.MacroBrinkmanCollideNaturalConvection(f, A["rho"], *V("ux", "uy", "uz"), alpha, p["nu"], g, A["tem"], *V("qx", "qy", "qz"), kappa, *G, p["tem0"], True, gsnap)
            f.Stream(); g.Stream()
            f.BoundaryCondition(P["f_wall"])
            pl.AD.BoundaryConditionSetT(g, P["tem"], *V("ux", "uy", "uz"), P["setT"])
            pl.AD.BoundaryConditionSetQ(g, P["qn"], *V("ux", "uy", "uz"), kappa, P["setQ"])
            g.BoundaryCondition(P["g_wall"])
            f.SmoothCorner(); g.SmoothCorner()
            swap(fpairs)
    else:
        def fargs(sw):
            names = dict(ux="uxp" if sw else "ux", uy="uyp" if sw else "uy", uz="uzp" if sw else "uz", qx="qxp" if sw else "qx", qy="qyp" if sw else "qy", qz="qzp" if sw else "qz")
            arrs = {k: A[v] for k, v in names.items() if d3 or not k.endswith("z")}
            ca = pl.collide_args(api.M_AD_BRINKMAN_NAT_CONV, True, p["nu"], gx=p["gx"], gy=p["gy"], gz=p["gz"], tem0=p["tem0"], rho=A["rho"], tem=A["tem"],
                                 alpha=alpha, diffusivity=kappa, snapshot=gsnap, **arrs)
            aux = pl.bc_aux(ux=arrs["ux"], uy=arrs["uy"], uz=arrs.get("uz"), diffusivity=kappa)
            return ca, aux
        (c0, a0), (c1, a1) = fargs(False), fargs(True)
        plan = pl.StepPlan(f, g).set_collide(c0, c1).set_stream(False)
        plan.add_bounce(f, P["f_wall"])
        plan.add_closure(g, api.BC_AD_SET_T, P["setT"], [P["tem"]], a0, a1)
        plan.add_closure(g, api.BC_AD_SET_Q, P["setQ"], [P["qn"]], a0, a1)
        plan.add_bounce(g, P["g_wall"])
        plan.set_smooth_corner(True, True).finalize()
        advance(plan, fpairs)
    res = {k: A[k].to_host() for k in FWD if d3 or not k.endswith("z")}
    res["gsnap"] = api.snapshot_to_host(g, gsnap)
    if not only_forward:
        pl.ANS.InitialCondition(f, *V("ux", "uy", "uz"), A["ip"], *V("iux", "iuy", "iuz"))
        pl.AAD.InitialCondition(g, *V("ux", "uy", "uz"), A["item"], *V("iqx", "iqy", "iqz"))
        if not fused:
            for _ in range(nt):
                pl.AAD.MacroBrinkmanCollideNaturalConvection(f, A["rho"], *V("ux", "uy", "uz"), A["ip"], *V("iux", "iuy", "iuz"), *V("imx", "imy", "imz"), alpha, p["nu"],
                                                             g, A["tem"], A["item"], *V("iqx", "iqy", "iqz"), kappa, *G, True, igsnap)
                f.iStream(); g.iStream()
                pl.AAD.iBoundaryConditionSetT(g, *V("ux", "uy", "uz"), P["setT"])
                pl.AAD.iBoundaryConditionSetQ(g, *V("ux", "uy", "uz"), P["setQ"])
                pl.AAD.iBoundaryConditionSetQ(g, *V("ux", "uy", "uz"), P["source"], 1.0)
                g.iBoundaryCondition(P["g_wall"])
                f.iBoundaryCondition(P["f_wall"])
                f.SmoothCorner(); g.SmoothCorner()
                swap(apairs)
        else:
            def aargs(sw):
                names = dict(iux="iuxp" if sw else "iux", iuy="iuyp" if sw else "iuy", iuz="iuzp" if sw else "iuz", iqx="iqxp" if sw else "iqx", iqy="iqyp" if sw else "iqy",
                             iqz="iqzp" if sw else "iqz")
                arrs = {k: A[v] for k, v in names.items() if d3 or not k.endswith("z")}
                fixed = {k: A[k] for k in ("rho", "ux", "uy", "uz", "tem", "ip", "imx", "imy", "imz", "item") if d3 or not k.endswith("z")}
                return pl.collide_args(api.M_AAD_NAT_CONV, True, p["nu"], gx=p["gx"], gy=p["gy"], gz=p["gz"], alpha=alpha, diffusivity=kappa, snapshot=igsnap, **fixed, **arrs)
            aux0 = pl.bc_aux(ux=A["ux"], uy=A["uy"], uz=A["uz"] if d3 else None)
            aux1 = pl.bc_aux(ux=A["ux"], uy=A["uy"], uz=A["uz"] if d3 else None, eps=1.0)
            plan = pl.StepPlan(f, g).set_collide(aargs(False), aargs(True)).set_stream(True)
            plan.add_closure(g, api.BC_AAD_ISET_T, P["setT"], [], aux0, aux0)
            plan.add_closure(g, api.BC_AAD_ISET_Q, P["setQ"], [], aux0, aux0)
            plan.add_closure(g, api.BC_AAD_ISET_Q, P["source"], [], aux1, aux1)
            plan.add_bounce(g, P["g_wall"], inverse=True)
            plan.add_bounce(f, P["f_wall"], inverse=True)
            plan.set_smooth_corner(True, True).finalize()
            advance(plan, apairs)
        dfdss = pl.DeviceArray(n, 0.0)
        pl.AAD.SensitivityTemperatureAtHeatSource(g, dfdss, *V("ux", "uy", "uz"), *V("imx", "imy", "imz"), dads, A["tem"], A["item"], *V("iqx", "iqy", "iqz"),
                                                  gsnap, igsnap, kappa, dkds, P["qn"], P["source"])
        res.update({k: A[k].to_host() for k in ADJ if d3 or not k.endswith("z")})
        res["igsnap"] = api.snapshot_to_host(g, igsnap)
        res["dfdss"] = dfdss.to_host()
    res["f.f0"], res["f.f"] = f.get_populations()
    res["g.f0"], res["g.f"] = g.get_populations()
    return res


def compare(a, b, what=""):
    assert set(a) == set(b), (sorted(a), sorted(b))
    for k in sorted(a):
        assert np.array_equal(a[k], b[k]), f"{what}: {k} differs (max abs {np.max(np.abs(a[k] - b[k])):.3e}, scale {np.max(np.abs(b[k])):.3e})"
