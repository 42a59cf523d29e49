"""GPU parity of the block-decomposed path (what replaces the reference's MPI build, d3q15.h:257-600 / 1306-1407):
all ranks of a PE grid live in ONE process on ONE device (pl_comm_init_loopback) and exchange their halos through the
same pack kernel, message layout and halo-aware pull the NCCL path uses.  Invariant: a decomposed run equals
the single-block run site for site — bit for bit when every block holds a multiple of 4 sites.  (Otherwise the last
nxyz%4 sites of a block take the reference's scalar-tail operation order, navierstokes_avx.h:180-200, in the decomposed run
and the AVX order in the single-block run — in the reference's MPI build as well — and agree to rounding only.)"""
import math

import numpy as np
import pytest

from helpers import random_pops

pytestmark = pytest.mark.gpu

C3 = np.array([[0, 1, 0, 0, -1, 0, 0, 1, -1, 1, 1, -1, 1, -1, -1],
               [0, 0, 1, 0, 0, -1, 0, 1, 1, -1, 1, -1, -1, 1, -1],
               [0, 0, 0, 1, 0, 0, -1, 1, 1, 1, -1, -1, -1, -1, 1]])
C2 = np.array([[0, 1, 0, -1, 0, 1, -1, -1, 1], [0, 0, 1, 0, -1, 1, 1, -1, -1], [0]*9])


@pytest.fixture
def world():
    """loopback communicator factory; everything created inside is freed before the communicator goes away"""
    import gc
    import panslbm2_b200 as pl
    made = []

    def start(n):
        pl.comm_init_loopback(n)
        made.append(n)
        return pl
    yield start
    gc.collect()
    if made:
        pl.comm_destroy()


def blocks(pl, dim, size, m):
    n = m[0]*m[1]*m[2]
    return [pl.D3Q15(*size, r, *m) if dim == 3 else pl.D2Q9(size[0], size[1], r, m[0], m[1]) for r in range(n)]


def to_global(lat, size, arrays):
    """assemble per-rank site arrays into the global site order"""
    lx, ly, lz = size
    out = np.zeros((lz, ly, lx))
    for l, a in zip(lat, arrays):
        out[l.offsetz:l.offsetz + l.nz, l.offsety:l.offsety + l.ny, l.offsetx:l.offsetx + l.nx] = np.asarray(a).reshape(l.nz, l.ny, l.nx)
    return out.reshape(-1)


def pops_global(lat, size, nc):
    """per-rank populations (reference layout) -> array [c][global site]"""
    outs = []
    per = [l.get_populations() for l in lat]
    for c in range(nc):
        outs.append(to_global(lat, size, [f0 if c == 0 else f.reshape(-1, nc - 1)[:, c - 1] for f0, f in per]))
    return np.stack(outs)


def scatter_pops(lat, size, P):
    """P[c][k][j][i] -> per-rank populations in the reference layout"""
    for l in lat:
        blk = P[:, l.offsetz:l.offsetz + l.nz, l.offsety:l.offsety + l.ny, l.offsetx:l.offsetx + l.nx].reshape(P.shape[0], -1)
        l.set_populations(np.ascontiguousarray(blk[0]), np.ascontiguousarray(blk[1:].T.reshape(-1)))


@pytest.mark.parametrize("dim,size,m", [(3, (8, 6, 6), (2, 1, 1)), (3, (7, 6, 5), (1, 2, 1)), (3, (6, 5, 9), (1, 1, 2)), (3, (9, 8, 7), (2, 2, 2)),
                                        (3, (11, 7, 9), (3, 2, 2)), (3, (8, 9, 4), (2, 3, 1)), (2, (9, 8, 1), (2, 2, 1)), (2, (10, 7, 1), (3, 1, 1)),
                                        (2, (7, 9, 1), (1, 2, 1))])
@pytest.mark.parametrize("inverse", [0, 1])
def test_stream_decomposed_equals_global_periodic_stream(world, dim, size, m, inverse):
    pl = world(m[0]*m[1]*m[2])
    nc = 15 if dim == 3 else 9
    CC = C3 if dim == 3 else C2
    lx, ly, lz = size
    f0, f = random_pops(lx*ly*lz, nc, 11)
    P = np.concatenate([f0[None, :], f.reshape(-1, nc - 1).T]).reshape(nc, lz, ly, lx)
    lat = blocks(pl, dim, size, m)
    scatter_pops(lat, size, P)
    want = P
    for _ in range(3):
        for l in lat:
            l.iStream() if inverse else l.Stream()
        s = -1 if inverse else 1
        want = np.stack([np.roll(want[c], (s*CC[2][c], s*CC[1][c], s*CC[0][c]), axis=(0, 1, 2)) for c in range(nc)])
    got = pops_global(lat, size, nc)
    assert np.array_equal(got, want.reshape(nc, -1))


def cavity(pl, api, dim, size, m, nt, fused):
    """test/cavityflow3D.cpp:32-59 / test/cavityflow.cpp:31-66 on every rank of the PE grid, ranks advanced in lockstep"""
    lx, ly, lz = size
    nu, u0, theta = 0.1, 0.1, 90.0
    lat = blocks(pl, dim, size, m)
    if dim == 3:
        wall = lambda i, j, k: np.where((i == 0) | (i == lx - 1) | (j == 0) | (j == ly - 1) | (k == 0), 1, 0)
        lid = lambda i, j, k: k == lz - 1
        uvals = [lambda i, j, k: u0*math.cos(theta*math.pi/180.0), lambda i, j, k: u0*math.sin(theta*math.pi/180.0), lambda i, j, k: 0.0]
    else:
        wall = lambda i, j: np.where((i == 0) | (i == lx - 1) | (j == 0), 1, 0)
        lid = lambda i, j: j == ly - 1
        uvals = [lambda i, j: u0, lambda i, j: 0.0]
    rho = [pl.DeviceArray(l.nxyz, 1.0) for l in lat]
    u = [[pl.DeviceArray(l.nxyz, 0.0) for _ in range(dim)] for l in lat]
    for l, r, v in zip(lat, rho, u):
        pl.NS.InitialCondition(l, r, *v)
    if not fused:
        for _ in range(nt):
            for l, r, v in zip(lat, rho, u):
                pl.NS.MacroCollide(l, r, *v, nu, True)
            for l in lat:
                l.Stream()
            for l in lat:
                l.BoundaryCondition(wall)
                pl.NS.BoundaryConditionSetU(l, *uvals, lid)
                l.SmoothCorner()
    else:
        names = ["ux", "uy", "uz"][:dim]
        plans = []
        for l, r, v in zip(lat, rho, u):
            plan = pl.StepPlan(l)
            plan.set_collide(pl.collide_args(api.M_NS_COLLIDE, True, nu, rho=r, **dict(zip(names, v))))
            plan.add_bounce(l, wall)
            plan.add_closure(l, api.BC_NS_SET_U, lid, uvals)
            plans.append(plan.set_smooth_corner(True).finalize())
        for _ in range(nt):
            for plan in plans:
                plan.advance(1, end_streamed=False)
        for plan in plans:
            plan.advance(0, end_streamed=True)
    nc = 15 if dim == 3 else 9
    res = {"rho": to_global(lat, size, [r.to_host() for r in rho])}
    for d in range(dim):
        res["u%d" % d] = to_global(lat, size, [v[d].to_host() for v in u])
    res["pops"] = pops_global(lat, size, nc)
    return res


@pytest.mark.parametrize("dim,size,m,nt", [(3, (12, 10, 8), (2, 2, 2), 24), (3, (13, 9, 8), (3, 1, 2), 15), (3, (16, 8, 8), (2, 1, 1), 12),
                                           (2, (16, 12, 1), (2, 2, 1), 30)])
@pytest.mark.parametrize("fused", [False, True])
def test_cavity_decomposed_equals_single_block(world, dim, size, m, nt, fused):
    from panslbm2_b200 import api
    pl = world(m[0]*m[1]*m[2])
    got = cavity(pl, api, dim, size, m, nt, fused)
    pl.comm_destroy(); pl.comm_init_loopback(1)
    want = cavity(pl, api, dim, size, (1, 1, 1), nt, fused=False)
    for k in want:
        assert np.array_equal(got[k], want[k]), f"{k} differs (max abs {np.max(np.abs(got[k] - want[k])):.3e})"


@pytest.mark.parametrize("fused", [False, True])
def test_cavity_decomposed_with_scalar_tail_blocks_agrees_to_rounding(world, fused):
    from panslbm2_b200 import api
    size, m, nt = (12, 10, 9), (2, 2, 2), 24          # blocks of 6x5x5 = 150 sites: 2 scalar-tail sites each
    pl = world(8)
    got = cavity(pl, api, 3, size, m, nt, fused)
    pl.comm_destroy(); pl.comm_init_loopback(1)
    want = cavity(pl, api, 3, size, (1, 1, 1), nt, fused=False)
    for k in want:
        assert np.max(np.abs(got[k] - want[k])) <= 1e-13*np.max(np.abs(want[k])), k


def heatsink(pl, api, size, m, nt):
    """forward + adjoint loops + sensitivity of production/heatsink3D.cpp through the fused plans (bench.HeatsinkSweep), every
    rank of the PE grid advanced in lockstep; fields assembled in global site order"""
    import bench
    import heatsink_case as H
    n = m[0]*m[1]*m[2]
    sw = [bench.HeatsinkSweep(pl, api, size, r, m) for r in range(n)]
    for s in sw:
        s.upload_design()
        s.init_forward()
    for _ in range(nt):
        for s in sw:
            s.fplan.advance(1, end_streamed=False)
    for s in sw:
        s.fplan.advance(0, end_streamed=True)
    for s in sw:
        s.init_adjoint()
    for _ in range(nt):
        for s in sw:
            s.aplan.advance(1, end_streamed=False)
    for s in sw:
        s.aplan.advance(0, end_streamed=True)
    for s in sw:
        s.sensitivity()
    lat = [s.f for s in sw]
    res = {k: to_global(lat, size, [s.A[k].to_host() for s in sw]) for k in sw[0].A}    # both halves of every swapped pair
    res["dfdss"] = to_global(lat, size, [s.dfdss.to_host() for s in sw])
    res["fpops"] = pops_global(lat, size, 15)
    res["gpops"] = pops_global([s.g for s in sw], size, 15)
    return res


@pytest.mark.parametrize("size,m,nt", [((12, 12, 8), (2, 2, 2), 14), ((12, 10, 12), (2, 1, 3), 9)])
def test_heatsink3d_decomposed_equals_single_block(world, size, m, nt):
    from panslbm2_b200 import api
    pl = world(m[0]*m[1]*m[2])
    got = heatsink(pl, api, size, m, nt)
    pl.comm_destroy(); pl.comm_init_loopback(1)
    want = heatsink(pl, api, size, (1, 1, 1), nt)
    for k in want:
        assert np.array_equal(got[k], want[k]), f"{k} differs (max abs {np.max(np.abs(got[k] - want[k])):.3e})"


def test_decomposed_lattice_without_communicator_fails_loudly():
    import panslbm2_b200 as pl
    l = pl.D3Q15(8, 8, 8, 0, 2, 1, 1)
    with pytest.raises(pl.PanslbmError):
        l.Stream()
