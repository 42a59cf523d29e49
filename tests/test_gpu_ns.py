"""GPU parity (through the C-ABI): particle ops + NS equations, CUDA vs the C oracle (and vs the reference build where
oracle/_ref travelled), bit-exact.  Also the fused plan vs call-by-call, and the committed golden fixtures."""
import hashlib
import os

import numpy as np
import pytest

from helpers import gcoords, i32, random_field, random_pops, same
from oracle import oracle as O

pytestmark = pytest.mark.gpu

SIZES = {2: [(8, 6, 1), (7, 5, 1), (9, 6, 1), (11, 5, 1), (37, 19, 1)], 3: [(6, 4, 4), (5, 3, 3), (7, 3, 3), (5, 5, 3), (19, 11, 7)]}
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def backends(dim):
    from cuda_ops import CudaOps
    bes = [O.Backend("orc", dim), CudaOps(dim)]
    if O.have_ref(dim):
        bes.append(O.Backend("ref", dim))
    return bes


def run_all(dim, size, fn, peid=0, m=(1, 1, 1), seed=1):
    """run fn(backend, lattice) on every backend from the same random populations; return list of results"""
    res = []
    for be in backends(dim):
        l = be.lattice(*size, peid, *m)
        f0, f = random_pops(l.nxyz, l.nc, seed)
        l.set(f0, f)
        extra = fn(be, l)
        res.append((be.kind, l.get(), extra))
        l.free()
    return res


def assert_all_same(res):
    k0, (a0, a), e0 = res[0]
    for k, (b0, b), e in res[1:]:
        assert same(a0, b0), (k0, k, "f0")
        assert same(a, b), (k0, k, "f")
        if e0 is not None:
            for x, y in zip(e0, e):
                assert same(x, y), (k0, k, "macro")


@pytest.mark.parametrize("dim", [2, 3])
def test_layout_roundtrip(dim):
    from cuda_ops import CudaOps
    for size in SIZES[dim]:
        l = CudaOps(dim).lattice(*size)
        f0, f = random_pops(l.nxyz, l.nc, 3)
        l.set(f0, f)
        g0, g = l.get()
        assert same(f0, g0) and same(f, g)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("inverse", [0, 1])
def test_stream(dim, inverse):
    for size in SIZES[dim]:
        def fn(be, l):
            for _ in range(3):
                (be.istream if inverse else be.stream)(l)
        assert_all_same(run_all(dim, size, fn))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("inverse", [0, 1])
def test_bounce(dim, inverse):
    for n, size in enumerate(SIZES[dim]):
        bct = i32(np.random.RandomState(5 + n).randint(0, 3, size=size[0]*size[1]*size[2]))
        assert_all_same(run_all(dim, size, lambda be, l: be.bc(l, bct, inverse), seed=2 + n))


@pytest.mark.parametrize("dim", [2, 3])
def test_bounce_interior_plane(dim):
    size = SIZES[dim][0]
    bct = i32(np.random.RandomState(11).randint(0, 2, size=size[0]*size[1]*size[2]))
    for axis in range(dim):
        for d in (-1, 1):
            for inverse in (0, 1):
                assert_all_same(run_all(dim, size, lambda be, l: be.bc_plane(l, axis, 2, d, bct, inverse), seed=9))


@pytest.mark.parametrize("dim", [2, 3])
def test_smooth_corner(dim):
    for size in SIZES[dim]:
        assert_all_same(run_all(dim, size, lambda be, l: be.smooth_corner(l), seed=3))


@pytest.mark.parametrize("dim", [2, 3])
def test_ns_init(dim):
    for size in SIZES[dim]:
        N = size[0]*size[1]*size[2]
        rho = random_field(N, 1, 0.9, 1.1); u = [random_field(N, 2 + d, -0.1, 0.1) for d in range(3)]
        assert_all_same(run_all(dim, size, lambda be, l: be.ns_init(l, rho, *u)))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("brinkman", [0, 1])
def test_ns_collide(dim, brinkman):
    for n, size in enumerate(SIZES[dim]):
        N = size[0]*size[1]*size[2]
        alpha = random_field(N, 7, 0.0, 50.0)

        def fn(be, l):
            m = [np.full(N, -7.0) for _ in range(4)]
            if brinkman:
                be.ns_macro_brinkman_collide(l, *m, 0.1, alpha, 1)
                be.ns_macro_brinkman_collide(l, *m, 0.03, alpha, 0)
            else:
                be.ns_macro_collide(l, *m, 0.1, 1)
                be.ns_macro_collide(l, *m, 0.02, 0)
            return m[:dim + 1]
        assert_all_same(run_all(dim, size, fn, seed=20 + n))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("kind", ["set_u", "set_rho"])
def test_ns_face_closures(dim, kind):
    for n, size in enumerate(SIZES[dim]):
        Gn = size[0]*size[1]*size[2]
        mask = i32(np.random.RandomState(3 + n).randint(0, 2, size=Gn))
        if kind == "set_u":
            v = [random_field(Gn, 50 + d, -0.1, 0.1) for d in range(3)]
        else:
            v = [random_field(Gn, 60, 0.95, 1.05), random_field(Gn, 61, -0.1, 0.1), random_field(Gn, 62, -0.1, 0.1)]
        assert_all_same(run_all(dim, size, lambda be, l: getattr(be, "ns_bc_" + kind)(l, *v, mask), seed=40 + n))


@pytest.mark.parametrize("dim", [2, 3])
def test_decomposed_block_closures(dim):
    size = (9, 8, 7 if dim == 3 else 1)
    m = (2, 2, 2 if dim == 3 else 1)
    Gn = size[0]*size[1]*size[2]
    i, j, k = gcoords(*size)
    bct = i32(np.where((i == 0) | (j == size[1] - 1), 1, np.where(k == 0, 2, 0)))
    lid = i32(j == size[1] - 1)
    v = [random_field(Gn, 70 + d, -0.1, 0.1) for d in range(3)]
    for peid in range(m[0]*m[1]*m[2]):
        def fn(be, l):
            be.bc(l, bct, 0)
            be.ns_bc_set_u(l, *v, lid)
            be.smooth_corner(l)
        assert_all_same(run_all(dim, size, fn, peid, m, seed=80 + peid))


def test_residual_normalize():
    from cuda_ops import CudaOps
    orc, cu = O.Backend("orc", 3), CudaOps(3)
    n = 100003
    u = [random_field(n, s, -1, 1) for s in range(6)]
    assert abs(cu.residual3(*u, n)/orc.residual3(*u, n) - 1) < 1e-13
    assert abs(cu.residual2(u[0], u[1], u[3], u[4], n)/orc.residual2(u[0], u[1], u[3], u[4], n) - 1) < 1e-13
    assert abs(cu.residual1(u[0], u[3], n)/orc.residual1(u[0], u[3], n) - 1) < 1e-13
    a, b = u[0].copy(), u[0].copy()
    orc.normalize(a, n); cu.normalize(b, n)
    assert same(a, b)


# ---------------------------------------------------------------------------------------------------------
def cavity_cuda(lx, ly, lz, nt, fused, dim=3):
    """test/cavityflow3D.cpp:32-59 (test/cavityflow.cpp:31-66 for dim 2) through the Python mirror of the reference API"""
    import math
    import panslbm2_b200 as pl
    from panslbm2_b200 import api
    nu, u0, theta = 0.1, 0.1, 90.0
    if dim == 3:
        pf = pl.D3Q15(lx, ly, lz)
        wall = lambda i, j, k: np.where((i == 0) | (i == lx - 1) | (j == 0) | (j == ly - 1) | (k == 0), 1, 0)
        lid = lambda i, j, k: k == lz - 1
        uvals = [lambda i, j, k: u0*math.cos(theta*math.pi/180.0), lambda i, j, k: u0*math.sin(theta*math.pi/180.0), lambda i, j, k: 0.0]
    else:
        pf = pl.D2Q9(lx, ly)
        wall = lambda i, j: np.where((i == 0) | (i == lx - 1) | (j == 0), 1, 0)
        lid = lambda i, j: j == ly - 1
        uvals = [lambda i, j: u0, lambda i, j: 0.0]
    N = pf.nxyz
    rho = pl.DeviceArray(N, 1.0)
    u = [pl.DeviceArray(N, 0.0) for _ in range(dim)]
    pl.NS.InitialCondition(pf, rho, *u)
    if not fused:
        for _ in range(nt):
            pl.NS.MacroCollide(pf, rho, *u, nu, True)
            pf.Stream()
            pf.BoundaryCondition(wall)
            pl.NS.BoundaryConditionSetU(pf, *uvals, lid)
            pf.SmoothCorner()
    else:
        names = ["ux", "uy", "uz"][:dim]
        plan = pl.StepPlan(pf)
        plan.set_collide(pl.collide_args(api.M_NS_COLLIDE, True, nu, rho=rho, **dict(zip(names, u))))
        plan.add_bounce(pf, wall)
        plan.add_closure(pf, api.BC_NS_SET_U, lid, uvals)
        plan.set_smooth_corner(True).finalize()
        # split the run to exercise both entry phases of pl_plan_advance
        first = nt//3
        plan.advance(first, end_streamed=False)
        plan.advance(nt - first, end_streamed=True)
    out = [rho.to_host()] + [a.to_host() for a in u]
    return out, pf.get_populations()


@pytest.mark.parametrize("shape", [(9, 8, 7, 40), (16, 16, 16, 25), (33, 9, 5, 30)])
def test_cavity3d_fused_equals_stepwise_equals_oracle(shape):
    lx, ly, lz, nt = shape
    a, pa = cavity_cuda(lx, ly, lz, nt, fused=False)
    b, pb = cavity_cuda(lx, ly, lz, nt, fused=True)
    for x, y in zip(a, b):
        assert same(x, y)
    assert same(pa[0], pb[0]) and same(pa[1], pb[1])
    m = [np.zeros(lx*ly*lz) for _ in range(4)]
    O.Backend("orc", 3).time_cavity3d(lx, ly, lz, nt, 0, *m)
    for x, y in zip(a, m):
        assert same(x, y)


def test_cavity2d_fused_equals_stepwise():
    a, pa = cavity_cuda(21, 17, 1, 60, fused=False, dim=2)
    b, pb = cavity_cuda(21, 17, 1, 60, fused=True, dim=2)
    for x, y in zip(a, b):
        assert same(x, y)
    assert same(pa[0], pb[0]) and same(pa[1], pb[1])
    assert np.max(np.abs(a[1])) > 1e-3


def digest(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def canon(a):
    return a + 0.0   # -0.0 -> +0.0 so that digests only see values


def test_golden_cavity3d_reference_fixture():
    """the reference's own committed configuration 31^3 / 1000 steps (test/cavityflow3D.cpp:32), fused path, vs the
    fixture generated by the reference build"""
    z = np.load(os.path.join(G, "cavity3d.npz"))
    lx, ly, lz, nt = [int(v) for v in z["a_shape"]]
    a, _ = cavity_cuda(lx, ly, lz, nt, fused=True)
    for name, x in zip(("rho", "ux", "uy", "uz"), a):
        assert same(x, z[f"a_{name}"]), name
    lx, ly, lz, nt = [int(v) for v in z["b_shape"]]
    b, _ = cavity_cuda(lx, ly, lz, nt, fused=True)
    for name, x in zip(("rho", "ux", "uy", "uz"), b):
        assert same(x[::37], z[f"b_{name}_s37"]), name


def test_golden_cavity3d_128_cube_200_steps():
    """SURVEY §8d cfg 3 parity case: test/cavityflow3D.cpp scaled to 128^3, 200 steps, fused in-place passes, against the
    fixture of the reference build: every 997th value and the SHA-256 of each whole field — bit for bit"""
    import hashlib
    z = np.load(os.path.join(G, "cavity3d.npz"))
    lx, ly, lz, nt = [int(v) for v in z["c_shape"]]
    assert (lx, ly, lz, nt) == (128, 128, 128, 200)
    c, _ = cavity_cuda(lx, ly, lz, nt, fused=True)
    for name, x in zip(("rho", "ux", "uy", "uz"), c):
        assert same(x[::997], z[f"c_{name}_s997"]), (name, float(np.max(np.abs(x[::997] - z[f"c_{name}_s997"]))))
        assert hashlib.sha256(np.ascontiguousarray(canon(x)).tobytes()).digest() == bytes(z[f"c_{name}_sha256"]), name
    assert np.max(np.abs(c[2])) > 1e-2      # the lid moves along y (theta = 90 degrees, cavityflow3D.cpp:33)


def cavity2d_host(be, lx, ly, nt, u0=0.1, nu=0.1):
    """test/cavityflow.cpp:31-66 call by call on a CPU backend (the C oracle / the reference build)"""
    l = be.lattice(lx, ly)
    n = lx*ly
    i, j, _ = gcoords(lx, ly, 1)
    rho, ux, uy, uz = np.ones(n), np.zeros(n), np.zeros(n), np.zeros(n)
    wall = i32(np.where((i == 0) | (i == lx - 1) | (j == 0), 1, 0))
    lid = i32(j == ly - 1)
    uxg, uyg, uzg = np.full(n, u0), np.zeros(n), np.zeros(n)
    be.ns_init(l, rho, ux, uy, uz)
    for _ in range(nt):
        be.ns_macro_collide(l, rho, ux, uy, uz, nu, 1)
        be.stream(l)
        be.bc(l, wall, 0)
        be.ns_bc_set_u(l, uxg, uyg, uzg, lid)
        be.smooth_corner(l)
    pops = l.get()
    l.free()
    return [rho, ux, uy], pops


def test_config0_cavityflow2d_101x101_10000_steps_equals_oracle():
    """BASELINE configs[0] / SURVEY §8d cfg 1: test/cavityflow.cpp as committed (D2Q9 101 x 101, nu = 0.1, u0 = 0.1), 10 000 steps:
    the fused CUDA plan against the C oracle (and the reference build where oracle/_ref travelled), bit for bit"""
    lx = ly = 101
    nt = 10000
    got, pg = cavity_cuda(lx, ly, 1, nt, fused=True, dim=2)
    checkers = [O.Backend("orc", 2)] + ([O.Backend("ref", 2)] if O.have_ref(2) else [])
    for be in checkers:
        want, pw = cavity2d_host(be, lx, ly, nt)
        for name, a, b in zip(("rho", "ux", "uy"), got, want):
            assert same(a, b), (be.kind, name, float(np.max(np.abs(a - b))))
        assert same(pg[0], pw[0]) and same(pg[1], pw[1]), be.kind
    assert np.max(np.abs(got[1])) > 1e-2
