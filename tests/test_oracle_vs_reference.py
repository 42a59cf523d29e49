"""Pins the plain-C restatement (oracle/lbm_oracle.c) bit-for-bit against the reference's own OpenMP+AVX
build (oracle/_ref, compiled from the unmodified headers).  CPU only.  Skipped where oracle/_ref is absent
(the committed fixtures in tests/golden then carry the pin, see test_oracle_golden.py)."""
import numpy as np
import pytest

from helpers import f64, gcoords, i32, random_field, random_pops, same
from oracle import oracle as O

DIMS = [d for d in (2, 3) if O.have_ref(d)]
pytestmark = pytest.mark.skipif(not DIMS, reason="oracle/_ref not built (no /root/reference here)")

# sizes chosen so nxyz % 4 covers 0..3 (AVX tail path) and x/y/z extents differ
SIZES = {2: [(8, 6, 1), (7, 5, 1), (9, 6, 1), (11, 5, 1)], 3: [(6, 4, 4), (5, 3, 3), (7, 3, 3), (5, 5, 3)]}


def pair(dim, size, peid=0, m=(1, 1, 1)):
    ref, orc = O.Backend("ref", dim), O.Backend("orc", dim)
    lr = ref.lattice(*size, peid, *m)
    lo = orc.lattice(*size, peid, *m)
    return ref, orc, lr, lo


def load(lr, lo, seed):
    f0, f = random_pops(lr.nxyz, lr.nc, seed)
    lr.set(f0, f)
    lo.set(f0, f)


def check(lr, lo):
    a0, a = lr.get()
    b0, b = lo.get()
    assert same(a0, b0) and same(a, b)


@pytest.mark.parametrize("dim", DIMS)
def test_decomposition_rule(dim):
    ref, orc = O.Backend("ref", dim), O.Backend("orc", dim)
    grids = [(1, 1, 1), (2, 1, 1), (2, 2, 1), (3, 2, 1)] + ([(2, 2, 2), (3, 2, 2)] if dim == 3 else [])
    for size in [(13, 11, 7 if dim == 3 else 1), (8, 8, 8 if dim == 3 else 1)]:
        for m in grids:
            for peid in range(m[0] * m[1] * m[2]):
                lr, lo = ref.lattice(*size, peid, *m), orc.lattice(*size, peid, *m)
                for k in ("nx", "ny", "nz", "nxyz", "offx", "offy", "offz", "pex", "pey", "pez", "nc"):
                    assert getattr(lr, k) == getattr(lo, k), (size, m, peid, k)
                lr.free(); lo.free()


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("inverse", [0, 1])
def test_stream(dim, inverse):
    for size in SIZES[dim]:
        ref, orc, lr, lo = pair(dim, size)
        load(lr, lo, 1)
        for _ in range(3):
            (ref.istream if inverse else ref.stream)(lr)
            (orc.istream if inverse else orc.stream)(lo)
        check(lr, lo)


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("inverse", [0, 1])
def test_bounce_all_faces(dim, inverse):
    for n, size in enumerate(SIZES[dim]):
        ref, orc, lr, lo = pair(dim, size)
        load(lr, lo, 2 + n)
        bct = i32(np.random.RandomState(5 + n).randint(0, 3, size=size[0] * size[1] * size[2]))
        ref.bc(lr, bct, inverse)
        orc.bc(lo, bct, inverse)
        check(lr, lo)


@pytest.mark.parametrize("dim", DIMS)
def test_bounce_interior_plane(dim):
    size = SIZES[dim][0]
    for axis in range(dim):
        for d in (-1, 1):
            for inverse in (0, 1):
                ref, orc, lr, lo = pair(dim, size)
                load(lr, lo, 9)
                bct = i32(np.random.RandomState(11).randint(0, 2, size=size[0] * size[1] * size[2]))  # BARRIER only
                ref.bc_plane(lr, axis, 2, d, bct, inverse)
                orc.bc_plane(lo, axis, 2, d, bct, inverse)
                check(lr, lo)


@pytest.mark.parametrize("dim", DIMS)
def test_smooth_corner(dim):
    for size in SIZES[dim]:
        ref, orc, lr, lo = pair(dim, size)
        load(lr, lo, 3)
        ref.smooth_corner(lr)
        orc.smooth_corner(lo)
        check(lr, lo)


@pytest.mark.parametrize("dim", DIMS)
def test_ns_init_and_collide(dim):
    for n, size in enumerate(SIZES[dim]):
        ref, orc, lr, lo = pair(dim, size)
        N = lr.nxyz
        rho = random_field(N, 1, 0.9, 1.1); ux = random_field(N, 2, -0.1, 0.1); uy = random_field(N, 3, -0.1, 0.1); uz = random_field(N, 4, -0.1, 0.1)
        ref.ns_init(lr, rho, ux, uy, uz)
        orc.ns_init(lo, rho, ux, uy, uz)
        check(lr, lo)
        load(lr, lo, 20 + n)
        outs = []
        for be, l in ((ref, lr), (orc, lo)):
            m = [np.full(N, -7.0) for _ in range(4)]
            be.ns_macro_collide(l, *m, 0.1, 1)
            be.ns_macro_collide(l, *m, 0.02, 0)
            outs.append(m)
        check(lr, lo)
        for a, b in zip(*outs):
            assert same(a, b)


@pytest.mark.parametrize("dim", DIMS)
def test_ns_brinkman_collide(dim):
    for n, size in enumerate(SIZES[dim]):
        ref, orc, lr, lo = pair(dim, size)
        N = lr.nxyz
        load(lr, lo, 30 + n)
        alpha = random_field(N, 7, 0.0, 50.0)
        outs = []
        for be, l in ((ref, lr), (orc, lo)):
            m = [np.full(N, -7.0) for _ in range(4)]
            be.ns_macro_brinkman_collide(l, *m, 0.1, alpha, 1)
            outs.append(m)
        check(lr, lo)
        for a, b in zip(*outs):
            assert same(a, b)


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("kind", ["set_u", "set_rho"])
def test_ns_face_closures(dim, kind):
    for n, size in enumerate(SIZES[dim]):
        ref, orc, lr, lo = pair(dim, size)
        G = size[0] * size[1] * size[2]
        load(lr, lo, 40 + n)
        mask = i32(np.random.RandomState(3 + n).randint(0, 2, size=G))
        if kind == "set_u":
            v = [random_field(G, 50 + d, -0.1, 0.1) for d in range(3)]
        else:
            v = [random_field(G, 60, 0.95, 1.05), random_field(G, 61, -0.1, 0.1), random_field(G, 62, -0.1, 0.1)]
        getattr(ref, "ns_bc_" + kind)(lr, *v, mask)
        getattr(orc, "ns_bc_" + kind)(lo, *v, mask)
        check(lr, lo)


@pytest.mark.parametrize("dim", DIMS)
def test_decomposed_block_ops(dim):
    """A rank's block (PEid > 0) sees global predicates through its offsets."""
    size = (9, 8, 7 if dim == 3 else 1)
    m = (2, 2, 2 if dim == 3 else 1)
    G = size[0] * size[1] * size[2]
    i, j, k = gcoords(*size)
    bct = i32(np.where((i == 0) | (j == size[1] - 1), 1, np.where(k == 0, 2, 0)))
    lid = i32(j == size[1] - 1)
    v = [random_field(G, 70 + d, -0.1, 0.1) for d in range(3)]
    for peid in range(m[0] * m[1] * m[2]):
        ref, orc, lr, lo = pair(dim, size, peid, m)
        load(lr, lo, 80 + peid)
        for be, l in ((ref, lr), (orc, lo)):
            be.bc(l, bct, 0)
            be.ns_bc_set_u(l, *v, lid)
            be.smooth_corner(l)
        check(lr, lo)


def test_cavity3d_sequence():
    """test/cavityflow3D.cpp call sequence, 9x8x7 box, 40 steps: every field bit-identical."""
    if 3 not in DIMS:
        pytest.skip("no 3-D reference build")
    ref, orc = O.Backend("ref", 3), O.Backend("orc", 3)
    lx, ly, lz = 9, 8, 7
    N = lx * ly * lz
    a = [np.zeros(N) for _ in range(4)]
    b = [np.zeros(N) for _ in range(4)]
    ref.time_cavity3d(lx, ly, lz, 40, 0, *a)
    orc.time_cavity3d(lx, ly, lz, 40, 0, *b)
    for x, y in zip(a, b):
        assert same(x, y)
    assert np.max(np.abs(a[1])) > 1e-3  # the lid actually drives a flow


def test_residual_and_normalize():
    dim = DIMS[0]
    ref, orc = O.Backend("ref", dim), O.Backend("orc", dim)
    n = 1001
    u = [random_field(n, s, -1, 1) for s in range(6)]
    assert ref.residual3(*u, n) == orc.residual3(*u, n)
    assert ref.residual2(u[0], u[1], u[3], u[4], n) == orc.residual2(u[0], u[1], u[3], u[4], n)
    assert ref.residual1(u[0], u[3], n) == orc.residual1(u[0], u[3], n)
    a, b = u[0].copy(), u[0].copy()
    ref.normalize(a, n)
    orc.normalize(b, n)
    assert same(a, b)


# ---------------------------------------------------------------------------------------------------------
# every collide model, closure, initial condition and sensitivity of NS / AD / ANS / AAD (tests/scenarios.py drives every
# backend the same way: random populations and fields, sizes covering every scalar-tail length)
import scenarios as S  # noqa: E402


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("model", S.FORWARD_MODELS + S.ADJOINT_MODELS)
def test_every_collide_model(dim, model):
    if model == "aad_natural_convection_massflow" and dim == 3:
        pytest.skip("the reference's D3Q15 overload does not compile")
    ref, orc = O.Backend("ref", dim), O.Backend("orc", dim)
    for n, size in enumerate(SIZES[dim]):
        S.assert_same(S.collide(ref, dim, model, size, 5 + n), S.collide(orc, dim, model, size, 5 + n), f"{model} {size}")


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("kind", S.CLOSURES)
def test_every_closure(dim, kind):
    if kind == "aad_iset_rho" and dim == 3:
        pytest.skip("the reference's D3Q15 overload does not compile")
    ref, orc = O.Backend("ref", dim), O.Backend("orc", dim)
    for n, size in enumerate(SIZES[dim]):
        S.assert_same(S.closure(ref, dim, kind, size, 9 + n), S.closure(orc, dim, kind, size, 9 + n), f"{kind} {size}")


@pytest.mark.parametrize("dim", DIMS)
def test_initial_conditions_and_sensitivities(dim):
    ref, orc = O.Backend("ref", dim), O.Backend("orc", dim)
    for n, size in enumerate(SIZES[dim]):
        S.assert_same(S.inits(ref, dim, size, 3 + n), S.inits(orc, dim, size, 3 + n), f"inits {size}")
        for kind in S.SENSITIVITIES:
            S.assert_same(S.sensitivity(ref, dim, kind, size, 11 + n), S.sensitivity(orc, dim, kind, size, 11 + n), f"{kind} {size}")


@pytest.mark.parametrize("dim", DIMS)
def test_heatsink_iteration_sequence(dim):
    """forward loop, adjoint loop and sensitivity of the heatsink drivers, op by op: every field bit-identical"""
    import heatsink_case as H
    size, nt = ((13, 17, 9), 25) if dim == 3 else ((23, 29, 1), 25)
    a = H.run_oplevel(O.Backend("ref", dim), dim, size, nt)
    b = H.run_oplevel(O.Backend("orc", dim), dim, size, nt)
    assert sorted(a) == sorted(b)
    for k in a:
        assert same(a[k], b[k]), k


# ---------------------------------------------------------------------------------------------------------
# NSin (src/equation/nsincompressible.h, D2Q9 only): the C restatement against the reference headers
@pytest.mark.skipif(not O.have_ref(2), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("size", [(8, 6, 1), (7, 5, 1), (9, 6, 1), (11, 5, 1), (23, 17, 1)])
def test_nsin_oracle_equals_reference(size):
    import scenarios as S
    ref, orc = O.Backend("ref", 2), O.Backend("orc", 2)
    if not ref.has("nsin_macro_collide"):
        pytest.skip("oracle/_ref predates the NSin entry points: make -C oracle ref")
    S.assert_same(S.nsin(ref, size, 4), S.nsin(orc, size, 4), f"NSin {size}")
