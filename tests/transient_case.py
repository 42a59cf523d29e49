"""The transient heatsink loops (BASELINE configs[4]; production/heatsink3D_transient.cpp:145-232 as restated in
tests/dropin/transient_dump.cpp) through the Python mirror of the device-pointer API, on top of the checkpoint-recompute state
store (panslbm2_b200/transient.py).  every = 1 keeps every step, which is what the reference does; every = K keeps every K-th
step plus a ring of K - 1 and recomputes the rest during the adjoint loop.  Returns the arrays tests/golden/transient.npz holds
(generated from the reference headers), so both modes are compared with the reference itself."""
import numpy as np

import heatsink_case as H


def run_transient_cuda(size, nt, every, stats=None):
    import panslbm2_b200 as pl
    from panslbm2_b200 import api
    from panslbm2_b200.transient import CheckpointedSweep
    p = H.params(3, size)
    P = H.predicates(p)
    f, g = pl.D3Q15(*size), pl.D3Q15(*size)          # forward lattices (also the ones segments are recomputed on)
    af, ag = pl.D3Q15(*size), pl.D3Q15(*size)        # adjoint lattices
    n = f.nxyz

    class _L:
        nx, ny, nz, offx, offy, offz = f.nx, f.ny, f.nz, 0, 0, 0
    alpha, kappa, dads, dkds = [pl.DeviceArray.from_host(np.ascontiguousarray(a)) for a in H.design_fields(p, *H.local_coords(_L))]
    names = ["rho", "ux", "uy", "uz", "tem", "qx", "qy", "qz"]
    T = nt - 1                                        # forward steps 1 .. nt-1 (heatsink3D_transient.cpp:150)

    def make_state():
        s = {k: pl.DeviceArray(n, 0.0) for k in names}
        s["gi"] = pl.DeviceArray(n*15, 0.0)
        return s
    s0 = make_state()
    s0["rho"].fill(1.0)
    pl.NS.InitialCondition(f, s0["rho"], s0["ux"], s0["uy"], s0["uz"])
    pl.AD.InitialCondition(g, s0["tem"], s0["ux"], s0["uy"], s0["uz"])

    def cargs(s):
        return pl.collide_args(api.M_AD_BRINKMAN_NAT_CONV, True, p["nu"], gx=p["gx"], gy=p["gy"], gz=p["gz"], tem0=p["tem0"], alpha=alpha, diffusivity=kappa,
                               snapshot=s["gi"], **{k: s[k] for k in names})

    def aux(s, eps=0.0):
        return pl.bc_aux(ux=s["ux"], uy=s["uy"], uz=s["uz"], diffusivity=kappa, eps=eps)
    plan = pl.StepPlan(f, g).set_collide(cargs(s0)).set_stream(False)
    plan.add_bounce(f, P["f_wall"])
    plan.add_closure(g, api.BC_AD_SET_T, P["setT"], [P["tem"]], aux(s0), aux(s0))
    plan.add_closure(g, api.BC_AD_SET_Q, P["setQ"], [P["qn"]], aux(s0), aux(s0))
    plan.add_bounce(g, P["g_wall"])
    plan.set_smooth_corner(True, True).finalize()
    sweep = CheckpointedSweep(plan, [f, g], T, every, make_state, lambda k, s: plan.rebind(k, collide=cargs(s), aux=[aux(s), aux(s)]), state0=s0)
    import time
    pl.synchronize(); t0 = time.perf_counter()
    sweep.forward()
    pl.synchronize(); t1 = time.perf_counter()

    # samples of three stored steps, taken while they are resident (a checkpointed sweep overwrites ring states later)
    res = {}
    tq = [1, nt//2, nt - 1]

    import math
    Lp = int(math.ceil(p["L"]))
    objective = [0.0]

    def sample(t, s):
        # objective: heat-patch temperature summed over every stored step (heatsink3D_transient.cpp:221-231) — taken on the device
        # while the state is resident (pl_reduce_box_sum); the reference sums the same values on the host in another order
        objective[0] += pl.box_sum(g, s["tem"], 0, Lp, 0, 1, 0, Lp)
        for q, tt in enumerate(tq):
            if tt == t:
                for k in ("rho", "ux", "uz", "tem", "qy"):
                    res[f"{k}@{q}"] = s[k].to_host()

    A = {k: pl.DeviceArray(n, 0.0) for k in H.ADJ}
    igi = pl.DeviceArray(n*15, 0.0)
    dfdss = pl.DeviceArray(n, 0.0)
    sT = sweep.state(T)
    sample(T, sT)
    pl.ANS.InitialCondition(af, sT["ux"], sT["uy"], sT["uz"], A["ip"], A["iux"], A["iuy"], A["iuz"])
    pl.AAD.InitialCondition(ag, sT["ux"], sT["uy"], sT["uz"], A["item"], A["iqx"], A["iqy"], A["iqz"])

    def aargs(s):
        return pl.collide_args(api.M_AAD_NAT_CONV, True, p["nu"], gx=p["gx"], gy=p["gy"], gz=p["gz"], alpha=alpha, diffusivity=kappa, snapshot=igi,
                               rho=s["rho"], ux=s["ux"], uy=s["uy"], uz=s["uz"], tem=s["tem"], **A)
    aplan = pl.StepPlan(af, ag).set_collide(aargs(sT)).set_stream(True)
    aplan.add_closure(ag, api.BC_AAD_ISET_T, P["setT"], [], aux(sT), aux(sT))
    aplan.add_closure(ag, api.BC_AAD_ISET_Q, P["setQ"], [], aux(sT), aux(sT))
    aplan.add_closure(ag, api.BC_AAD_ISET_Q, P["source"], [], aux(sT, 1.0), aux(sT, 1.0))
    aplan.add_bounce(ag, P["g_wall"], inverse=True)
    aplan.add_bounce(af, P["f_wall"], inverse=True)
    aplan.set_smooth_corner(True, True).finalize()

    def visit(t, s, s_next):
        sample(t, s)
        aplan.rebind(aplan.next_set(), collide=aargs(s), aux=[aux(s), aux(s), aux(s, 1.0)])
        aplan.advance(1, end_streamed=False)
        pl.AAD.SensitivityTemperatureAtHeatSource(ag, dfdss, s["ux"], s["uy"], s["uz"], A["imx"], A["imy"], A["imz"], dads, s["tem"], A["item"],
                                                  A["iqx"], A["iqy"], A["iqz"], s["gi"], igi, kappa, dkds, P["qn"], P["source"])
    pl.synchronize(); t2 = time.perf_counter()
    sweep.backward(visit, t_hi=T - 1, t_lo=0)          # for (t = nt - 2; t >= 0; --t)   heatsink3D_transient.cpp:190
    pl.synchronize(); t3 = time.perf_counter()
    # the loop body ends with iStream + closures + SmoothCorner of the last visit
    aplan.advance(0, end_streamed=True)
    res.update({k: A[k].to_host() for k in H.ADJ})
    res["dfdss"] = dfdss.to_host()
    res["extra"] = np.array([objective[0]])
    res["f.f0"], res["f.f"] = af.get_populations()
    res["g.f0"], res["g.f"] = ag.get_populations()
    if stats is not None:
        stats["recomputed"] = sweep.recomputed
        stats["states"] = len(sweep.perm) + len(sweep.ring)
        stats["forward_ms_per_step"] = 1e3*(t1 - t0)/T
        stats["adjoint_ms_per_step"] = 1e3*(t3 - t2)/T          # adjoint + sensitivity + the recomputed forward steps
        stats["state_bytes"] = stats["states"]*23*n*8
        stats["checkpoint_bytes"] = len(sweep.cps)*2*15*n*8
    sweep.free()
    return res
