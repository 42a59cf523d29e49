"""pl_plan_rebind through the device-pointer API: a loop whose arrays change every step (the transient drivers keep one set of
macroscopic arrays per time step, production/heatsink3D_transient.cpp:156-176) advanced as fused passes with the plan re-bound
per step must equal the same loop issued call by call — macroscopic fields of every step, populations at the end.
Two cases: NS cavity (collide arguments re-bound) and the thermal heatsink forward loop (collide arguments AND the velocity fields
the SetT/SetQ closures read)."""
import numpy as np
import pytest

import heatsink_case as H

pytestmark = pytest.mark.gpu


def test_rebind_collide_arguments_every_step():
    import math
    import panslbm2_b200 as pl
    from panslbm2_b200 import api
    lx, ly, lz, nt = 14, 12, 10, 17
    nu, u0 = 0.1, 0.1
    wall = lambda i, j, k: np.where((i == 0) | (i == lx - 1) | (j == 0) | (j == ly - 1) | (k == 0), 1, 0)
    lid = lambda i, j, k: k == lz - 1
    uvals = [lambda i, j, k: 0.0*u0, lambda i, j, k: u0, lambda i, j, k: 0.0]

    def run(fused):
        pf = pl.D3Q15(lx, ly, lz)
        N = pf.nxyz
        rho = [pl.DeviceArray(N, 1.0) for _ in range(nt + 1)]
        u = [[pl.DeviceArray(N, 0.0) for _ in range(3)] for _ in range(nt + 1)]
        pl.NS.InitialCondition(pf, rho[0], *u[0])
        args = lambda t: pl.collide_args(api.M_NS_COLLIDE, True, nu, rho=rho[t], ux=u[t][0], uy=u[t][1], uz=u[t][2])
        if not fused:
            for t in range(1, nt + 1):
                pl.NS.MacroCollide(pf, rho[t], *u[t], nu, True)
                pf.Stream()
                pf.BoundaryCondition(wall)
                pl.NS.BoundaryConditionSetU(pf, *uvals, lid)
                pf.SmoothCorner()
        else:
            plan = pl.StepPlan(pf).set_collide(args(1))
            plan.add_bounce(pf, wall).add_closure(pf, api.BC_NS_SET_U, lid, uvals).set_smooth_corner(True).finalize()
            for t in range(1, nt + 1):
                plan.rebind(plan.next_set(), collide=args(t))
                plan.advance(1, end_streamed=(t == nt))
        return [r.to_host() for r in rho], [[a.to_host() for a in ut] for ut in u], pf.get_populations()

    ra, ua, pa = run(False)
    rb, ub, pb = run(True)
    for t in range(1, nt + 1):
        assert np.array_equal(ra[t], rb[t]), f"rho[{t}]"
        for d in range(3):
            assert np.array_equal(ua[t][d], ub[t][d]), f"u[{t}][{d}]"
    assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1])
    assert math.isfinite(float(rb[nt].sum()))


def test_rebind_closure_fields_every_step():
    import panslbm2_b200 as pl
    from panslbm2_b200 import api
    size, nt = (16, 14, 12), 13
    p = H.params(3, size)
    P = H.predicates(p)

    def run(fused):
        f, g = pl.D3Q15(*size), pl.D3Q15(*size)
        N = f.nxyz

        class _L:
            nx, ny, nz, offx, offy, offz = f.nx, f.ny, f.nz, 0, 0, 0
        alpha, kappa, _, _ = [pl.DeviceArray.from_host(np.ascontiguousarray(a)) for a in H.design_fields(p, *H.local_coords(_L))]
        names = ["rho", "ux", "uy", "uz", "tem", "qx", "qy", "qz"]
        A = [{k: pl.DeviceArray(N, 1.0 if k == "rho" else 0.0) for k in names} for _ in range(nt + 1)]
        snap = [pl.DeviceArray(N*15) for _ in range(nt + 1)]
        pl.NS.InitialCondition(f, A[0]["rho"], A[0]["ux"], A[0]["uy"], A[0]["uz"])
        pl.AD.InitialCondition(g, A[0]["tem"], A[0]["ux"], A[0]["uy"], A[0]["uz"])
        cargs = lambda t: pl.collide_args(api.M_AD_BRINKMAN_NAT_CONV, True, p["nu"], gx=p["gx"], gy=p["gy"], gz=p["gz"], tem0=p["tem0"], alpha=alpha,
                                          diffusivity=kappa, snapshot=snap[t], **A[t])
        aux = lambda t: pl.bc_aux(ux=A[t]["ux"], uy=A[t]["uy"], uz=A[t]["uz"], diffusivity=kappa)
        if not fused:
            for t in range(1, nt + 1):
                a = A[t]
                pl.AD.MacroBrinkmanCollideNaturalConvection(f, a["rho"], a["ux"], a["uy"], a["uz"], alpha, p["nu"], g, a["tem"], a["qx"], a["qy"], a["qz"], kappa,
                                                            p["gx"], p["gy"], p["gz"], p["tem0"], True, snap[t])
                f.Stream(); g.Stream()
                f.BoundaryCondition(P["f_wall"])
                pl.AD.BoundaryConditionSetT(g, P["tem"], a["ux"], a["uy"], a["uz"], P["setT"])
                pl.AD.BoundaryConditionSetQ(g, P["qn"], a["ux"], a["uy"], a["uz"], kappa, P["setQ"])
                g.BoundaryCondition(P["g_wall"])
                f.SmoothCorner(); g.SmoothCorner()
        else:
            plan = pl.StepPlan(f, g).set_collide(cargs(1)).set_stream(False)
            plan.add_bounce(f, P["f_wall"])
            plan.add_closure(g, api.BC_AD_SET_T, P["setT"], [P["tem"]], aux(1), aux(1))
            plan.add_closure(g, api.BC_AD_SET_Q, P["setQ"], [P["qn"]], aux(1), aux(1))
            plan.add_bounce(g, P["g_wall"])
            plan.set_smooth_corner(True, True).finalize()
            for t in range(1, nt + 1):
                # the collide of step t and the closures that follow it read/write the arrays of step t: one argument set
                plan.rebind(plan.next_set(), collide=cargs(t), aux=[aux(t), aux(t)])
                plan.advance(1, end_streamed=(t == nt))
        out = {(t, k): A[t][k].to_host() for t in (1, nt//2, nt) for k in names}
        return out, f.get_populations(), g.get_populations()

    a, fa, ga = run(False)
    b, fb, gb = run(True)
    for key in a:
        assert np.array_equal(a[key], b[key]), key
    for x, y in zip(fa + ga, fb + gb):
        assert np.array_equal(x, y)
