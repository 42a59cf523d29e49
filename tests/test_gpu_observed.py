"""Save policy of the fused plan (pl_plan_advance_observed): the reference stores rho, u, T, q and the thermal snapshot at every
site on every step (production/heatsink3D.cpp:151, advection_avx.h:1040-1052) although nobody looks at them except Residual every
`dt` steps (:152-160) and the code after the loop (:227-246).  A plan advanced with save_last = 2 stores them at every site only
in the last two collides of a call (and on the closure planes, where the closures read the velocities, in all of them).  These
tests pin what a caller can see: at every point where it may look, every array is bit-identical to the store-every-step run."""
import numpy as np
import pytest

import heatsink_case as H

pytestmark = pytest.mark.gpu


def _observer(log):
    import panslbm2_b200 as pl

    def look(A, gsnap, igsnap):
        names = sorted(A)
        snap = {k: A[k].to_host() for k in names}
        snap["gsnap"], snap["igsnap"] = gsnap.to_host(), igsnap.to_host()
        n = A["ux"].n
        fw = [A[k] for k in ("ux", "uy", "uz") if k in A] + [A[k] for k in ("uxp", "uyp", "uzp") if k in A]
        snap["residual_u"] = np.nan_to_num(np.array([pl.Residual(*fw, n)]), nan=-1.0)      # 0/0 right after the first collide of a fluid at rest
        log.append(snap)
    return look


@pytest.mark.parametrize("dim,size,nt,chunks", [(3, (13, 12, 10), 23, (3, 4)), (3, (12, 9, 11), 16, (1, 5)), (2, (21, 17, 1), 29, (2, 6)), (3, (9, 7, 6), 12, (2, 1))])
def test_everything_a_caller_can_see_is_identical(dim, size, nt, chunks):
    a_log, b_log = [], []
    a = H.run_cuda(dim, size, nt, fused=True, chunks=chunks, observe=_observer(a_log))
    b = H.run_cuda(dim, size, nt, fused=True, chunks=chunks, save_last=2, observe=_observer(b_log))
    H.compare(b, a, "save_last=2 vs every step")
    assert len(a_log) == len(b_log) and len(a_log) >= 4
    for n, (x, y) in enumerate(zip(a_log, b_log)):
        # chunks of one collide leave the other argument set to the collide before them, which an earlier chunk stored
        H.compare(y, x, f"observation {n}")


def test_matches_the_reference_fixture_and_the_stepwise_run():
    from test_gpu_full import check_fixture, heatsink_cases
    dim, size, nt = heatsink_cases()["hs3d"]
    res = H.run_cuda(dim, size, nt, fused=True, chunks=(3, 7), save_last=2)
    check_fixture("hs3d", res)
    H.compare(res, H.run_cuda(dim, size, nt, fused=False), "save_last=2 vs call by call")


def _single_set_run(size, n1, n2, elide):
    """forward heatsink loop with ONE argument set (no std::swap): n1 storing collides, n2 collides that store (elide=False) or do
    not (save_last = 0), then two storing ones and the closing Stream"""
    import panslbm2_b200 as pl
    from panslbm2_b200 import api
    p = H.params(3, size)
    f, g = pl.D3Q15(*size), pl.D3Q15(*size)
    n = f.nxyz

    class _L:
        nx, ny, nz, offx, offy, offz = f.nx, f.ny, f.nz, 0, 0, 0
    alpha, kappa, _, _ = [pl.DeviceArray.from_host(a) for a in H.design_fields(p, *H.local_coords(_L))]
    P = H.predicates(p)
    names = ["rho", "ux", "uy", "uz", "tem", "qx", "qy", "qz"]
    A = {k: pl.DeviceArray(n, 0.0) for k in names}
    A["rho"].fill(1.0)
    gsnap = pl.DeviceArray(n*15, 0.0)
    pl.NS.InitialCondition(f, A["rho"], A["ux"], A["uy"], A["uz"])
    pl.AD.InitialCondition(g, A["tem"], A["ux"], A["uy"], A["uz"])
    ca = pl.collide_args(api.M_AD_BRINKMAN_NAT_CONV, True, p["nu"], gx=p["gx"], gy=p["gy"], gz=p["gz"], tem0=p["tem0"], alpha=alpha, diffusivity=kappa,
                         snapshot=gsnap, **A)
    aux = pl.bc_aux(ux=A["ux"], uy=A["uy"], uz=A["uz"], diffusivity=kappa)
    plan = pl.StepPlan(f, g).set_collide(ca, ca).set_stream(False)
    plan.add_bounce(f, P["f_wall"]).add_closure(g, api.BC_AD_SET_T, P["setT"], [P["tem"]], aux, aux)
    plan.add_closure(g, api.BC_AD_SET_Q, P["setQ"], [P["qn"]], aux, aux).add_bounce(g, P["g_wall"])
    plan.set_smooth_corner(True, True).finalize()
    grab = lambda: {**{k: A[k].to_host() for k in names}, "gsnap": gsnap.to_host()}
    plan.advance(n1, end_streamed=False)
    first = grab()
    plan.advance(n2, end_streamed=False, save_last=0 if elide else None)
    mid = grab()
    plan.advance(2, end_streamed=True, save_last=2)
    last = grab()
    last["f.f0"], last["f.f"] = f.get_populations()
    last["g.f0"], last["g.f"] = g.get_populations()
    return first, mid, last


def test_unobserved_steps_really_skip_their_stores():
    """save_last = 0: the interior of the saved fields still holds what the last STORING collide left (the elided passes do not
    write there), the closure planes — where SetT/SetQ read the velocities — are current, and the next storing collide leaves
    everything exactly as a run that stored all along"""
    from helpers import gcoords
    size, n1, n2 = (14, 12, 10), 4, 6
    f0, fm, fl = _single_set_run(size, n1, n2, elide=False)
    e0, em, el = _single_set_run(size, n1, n2, elide=True)
    H.compare(e0, f0, "before")
    H.compare(el, fl, "after the next storing collides")
    i, j, k = gcoords(*size)
    n = i.size
    # two sites in from every wall: the sites next to two walls form the SmoothCorner tubes, which the boundary pass owns
    interior = (i > 1) & (i < size[0] - 2) & (j > 1) & (j < size[1] - 2) & (k > 1) & (k < size[2] - 2)
    interior &= np.arange(n) < 4*(n//4)                       # the sites of the last incomplete AVX pack belong to the boundary pass
    for name in ("rho", "ux", "uy", "uz", "tem", "qx", "qy", "qz"):
        assert np.array_equal(em[name][interior], e0[name][interior]), f"{name}: an elided pass stored in the interior"
        assert not np.array_equal(fm[name][interior], f0[name][interior]), f"{name}: the storing run did not move (test too weak)"
    # the closure planes are current after every pass: the x walls belong to the interior kernel + k_xclose, the others to the boundary pass
    plane = (i == 0) | (i == size[0] - 1) | (j == 0) | (j == size[1] - 1) | (k == 0) | (k == size[2] - 1)
    for name in ("ux", "uy", "uz"):
        assert np.array_equal(em[name][plane], fm[name][plane]), f"{name} on the closure planes"
