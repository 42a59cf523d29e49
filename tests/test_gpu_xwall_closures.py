"""Inlet / outlet closures on the x walls inside fused plans (the channel of test/nssens.cpp, test/nssens3D.cpp and production/nsopt.cpp:
SetU on x = 0, SetRho on x = lx-1, bounce-back elsewhere; the adjoint loop with iSetU / iSetRho and iStream).  The x wall closures of
a plan run ahead of the fused pass (k_xclose) and hand their results to the interior kernel through the periodic wrap slots: the
fused run must equal the same loop issued call by call, whose every operation is pinned against the reference build."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def channel(dim, size, nt, fused, adjoint, partial=False):
    import panslbm2_b200 as pl
    from panslbm2_b200 import api
    lx, ly, lz = size
    d3 = dim == 3
    pf = pl.D3Q15(lx, ly, lz) if d3 else pl.D2Q9(lx, ly)
    n = pf.nxyz
    nu, u0 = 0.1, 0.03
    idx = np.arange(n)
    i, j, k = idx % lx, (idx//lx) % ly, idx//(lx*ly)
    alpha = pl.DeviceArray.from_host(np.where(((i - lx//2)**2 + (j - ly//2)**2 < (ly//5)**2), 0.6, 0.0).astype(np.float64))
    if d3:
        wall = lambda i, j, k: np.where((j == 0) | (j == ly - 1) | (k == 0) | (k == lz - 1), 1, 0)
        inlet = lambda i, j, k: (i == 0) & (j > 0) & (j < ly - 1) & (k > 0) & (k < lz - 1)
        outlet = lambda i, j, k: (i == lx - 1) & (j > 0) & (j < ly - 1) & (k > 0) & (k < lz - 1)
        uin = [lambda i, j, k: u0*(1.0 - ((2.0*j - (ly - 1))/(ly - 1))**2), lambda i, j, k: 0.0*j, lambda i, j, k: 0.0*j]
        rout = [lambda i, j, k: 1.0 + 0.0*j, lambda i, j, k: 0.0*j, lambda i, j, k: 0.0*j]
    else:
        wall = lambda i, j: np.where((j == 0) | (j == ly - 1), 1, 0)
        inlet = lambda i, j: (i == 0) & (j > 0) & (j < ly - 1)
        outlet = lambda i, j: (i == lx - 1) & (j > 0) & (j < ly - 1)
        uin = [lambda i, j: u0*(1.0 - ((2.0*j - (ly - 1))/(ly - 1))**2), lambda i, j: 0.0*j]
        rout = [lambda i, j: 1.0 + 0.0*j, lambda i, j: 0.0*j]
    if partial:
        # the closures cover only the middle of the x planes: the rest of each plane keeps the periodic wrap of Stream()
        full_in, full_out = inlet, outlet
        mid = lambda j: (j > ly//3) & (j < ly - 1 - ly//3)
        inlet = (lambda i, j, k: full_in(i, j, k) & mid(j)) if d3 else (lambda i, j: full_in(i, j) & mid(j))
        outlet = (lambda i, j, k: full_out(i, j, k) & mid(j)) if d3 else (lambda i, j: full_out(i, j) & mid(j))
    names = ["ux", "uy", "uz"][:dim]
    rho = pl.DeviceArray(n, 1.0)
    u = [pl.DeviceArray(n, 0.0) for _ in range(dim)]
    pl.NS.InitialCondition(pf, rho, *u)
    # forward loop (always, call by call or fused): the adjoint loop needs its fields
    if not fused or adjoint:
        for _ in range(nt):
            pl.NS.MacroBrinkmanCollide(pf, rho, *u, nu, alpha, True)
            pf.Stream()
            pf.BoundaryCondition(wall)
            pl.NS.BoundaryConditionSetU(pf, *uin, inlet)
            pl.NS.BoundaryConditionSetRho(pf, *rout, outlet)
    else:
        plan = pl.StepPlan(pf).set_collide(pl.collide_args(api.M_NS_BRINKMAN, True, nu, rho=rho, alpha=alpha, **dict(zip(names, u))))
        plan.add_bounce(pf, wall).add_closure(pf, api.BC_NS_SET_U, inlet, uin).add_closure(pf, api.BC_NS_SET_RHO, outlet, rout).finalize()
        plan.advance(nt//3, end_streamed=False)
        plan.advance(nt - nt//3, end_streamed=True)
    out = {"rho": rho.to_host(), **{nm: a.to_host() for nm, a in zip(names, u)}}
    if adjoint:
        A = {nm: pl.DeviceArray(n, 0.0) for nm in ["ip", "iux", "iuy", "iuz", "imx", "imy", "imz"]}
        iu, im = [A[k] for k in ["iux", "iuy", "iuz"][:dim]], [A[k] for k in ["imx", "imy", "imz"][:dim]]
        pl.ANS.InitialCondition(pf, *u, A["ip"], *iu)
        if not fused:
            for _ in range(nt):
                pl.ANS.MacroBrinkmanCollide(pf, rho, *u, A["ip"], *iu, *im, nu, alpha, True)
                pf.iStream()
                pf.iBoundaryCondition(wall)
                pl.ANS.iBoundaryConditionSetU(pf, *uin, inlet, 1.0)
                pl.ANS.iBoundaryConditionSetRho(pf, outlet)
        else:
            kw = dict(rho=rho, alpha=alpha, ip=A["ip"], **dict(zip(names, u)), **dict(zip(["iux", "iuy", "iuz"][:dim], iu)), **dict(zip(["imx", "imy", "imz"][:dim], im)))
            plan = pl.StepPlan(pf).set_collide(pl.collide_args(api.M_ANS_BRINKMAN, True, nu, **kw)).set_stream(True)
            plan.add_bounce(pf, wall, inverse=True)
            plan.add_closure(pf, api.BC_ANS_ISET_U, inlet, uin, pl.bc_aux(eps=1.0), pl.bc_aux(eps=1.0))
            plan.add_closure(pf, api.BC_ANS_ISET_RHO, outlet, [])
            plan.finalize()
            plan.advance(nt//2, end_streamed=False)
            plan.advance(nt - nt//2, end_streamed=True)
        out.update({k: A[k].to_host() for k in A if dim == 3 or not k.endswith("z")})
    f0, f = pf.get_populations()
    out["f0"], out["f"] = f0, f
    return out


@pytest.mark.parametrize("dim,size,nt", [(2, (26, 15, 1), 40), (3, (14, 11, 9), 30), (3, (13, 9, 7), 21)])
@pytest.mark.parametrize("adjoint", [False, True])
@pytest.mark.parametrize("partial", [False, True])
def test_fused_channel_equals_stepwise(dim, size, nt, adjoint, partial):
    a = channel(dim, size, nt, fused=False, adjoint=adjoint, partial=partial)
    b = channel(dim, size, nt, fused=True, adjoint=adjoint, partial=partial)
    assert sorted(a) == sorted(b)
    for k in sorted(a):
        assert np.array_equal(a[k], b[k]), f"{k}: max abs diff {np.max(np.abs(a[k] - b[k])):.3e}"
    assert np.max(np.abs(a["ux"])) > 1e-3
