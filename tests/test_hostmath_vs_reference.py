"""The arithmetic of the PRODUCT code on the CPU: panslbm2_b200/csrc/*.cuh site functions (every collide model,
boundary closure, InitialCondition and sensitivity formula the CUDA kernels execute per site), compiled for the host by
tests/hostmath, against the reference's own OpenMP+AVX build (oracle/_ref), bit for bit.  Skipped where oracle/_ref is
absent; the CUDA kernels themselves are checked by the -m gpu tests."""
import pytest

import scenarios as S
from helpers import hostmath_backend
from oracle import oracle as O

DIMS = [d for d in (2, 3) if O.have_ref(d)]
pytestmark = pytest.mark.skipif(not DIMS, reason="oracle/_ref not built (no /root/reference here)")

# nxyz % 4 covers 0..3 (AVX tail path); extents differ per axis
SIZES = {2: [(8, 6, 1), (7, 5, 1), (9, 6, 1), (11, 5, 1)], 3: [(6, 4, 4), (5, 3, 3), (7, 3, 3), (5, 5, 3)]}


def both(dim):
    return O.Backend("ref", dim), hostmath_backend(dim)


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("model", S.FORWARD_MODELS + S.ADJOINT_MODELS)
def test_collide_models(dim, model):
    if model.endswith("massflow") and dim == 3:
        pytest.skip("the reference's D3Q15 MassFlow overload does not compile")
    ref, hm = both(dim)
    for n, size in enumerate(SIZES[dim]):
        S.assert_same(S.collide(ref, dim, model, size, 3 + n), S.collide(hm, dim, model, size, 3 + n), f"{model} {size}")


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("kind", S.CLOSURES)
def test_closures(dim, kind):
    if kind == "aad_iset_rho" and dim == 3:
        pytest.skip("the reference's D3Q15 AAD::iBoundaryConditionSetRho does not compile")
    ref, hm = both(dim)
    for n, size in enumerate(SIZES[dim]):
        S.assert_same(S.closure(ref, dim, kind, size, 5 + n), S.closure(hm, dim, kind, size, 5 + n), f"{kind} {size}")


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("kind", S.SENSITIVITIES)
def test_sensitivities(dim, kind):
    ref, hm = both(dim)
    for n, size in enumerate(SIZES[dim]):
        S.assert_same(S.sensitivity(ref, dim, kind, size, 9 + n), S.sensitivity(hm, dim, kind, size, 9 + n), f"{kind} {size}")


@pytest.mark.parametrize("dim", DIMS)
def test_initial_conditions(dim):
    ref, hm = both(dim)
    for n, size in enumerate(SIZES[dim]):
        S.assert_same(S.inits(ref, dim, size, 2 + n), S.inits(hm, dim, size, 2 + n), f"init {size}")


@pytest.mark.parametrize("dim", DIMS)
def test_closures_on_decomposed_blocks(dim):
    size = (9, 8, 7 if dim == 3 else 1)
    m = (2, 2, 2 if dim == 3 else 1)
    ref, hm = both(dim)
    for peid in range(m[0]*m[1]*m[2]):
        for kind in ("ad_set_t", "ad_set_q_field", "aad_iset_q"):
            S.assert_same(S.closure(ref, dim, kind, size, 20 + peid, peid, m), S.closure(hm, dim, kind, size, 20 + peid, peid, m), f"{kind} pe{peid}")


def test_nsin_site_math():
    """NSin (nsincompressible.h, D2Q9 only): collide models 13 / 14, closures 12 / 13 and InitialCondition family 5 of the product code
    against the reference headers"""
    if 2 not in DIMS:
        pytest.skip("oracle/_ref (2-D) not built")
    ref, hm = both(2)
    if not ref.has("nsin_macro_collide"):
        pytest.skip("oracle/_ref predates the NSin entry points: make -C oracle ref")
    for n, size in enumerate(SIZES[2] + [(23, 17, 1)]):
        S.assert_same(S.nsin(ref, size, 4 + n, steps=0), S.nsin(hm, size, 4 + n, steps=0), f"NSin {size}")


# ---------------------------------------------------------------------------------------------------------
# pl_set_scalar_order: a caller built WITHOUT _USE_AVX_DEFINES (production/nsopt.cpp:2).  The product's site math with every site
# taken as a scalar-order site, against the reference headers compiled without the macro (oracle/_ref/*_scalar.so).
def _scalar_pair(dim):
    import ctypes as C
    if not O.have_ref_scalar(dim):
        pytest.skip("oracle/_ref/*_scalar.so not built (make -C oracle ref)")
    hm = hostmath_backend(dim)
    hm._fn("set_scalar_build")(C.c_int(1))
    return O.Backend("ref_scalar", dim), hm


def _ns_collides(be, dim, size, seed):
    from helpers import random_field, random_pops
    import numpy as np
    l = be.lattice(*size)
    n = l.nxyz
    l.set(*random_pops(n, l.nc, seed))
    alpha = random_field(n, seed + 5, 0.0, 40.0)
    res = []
    for k, (issave, nu, al) in enumerate(((1, 0.07, None), (0, 0.02, None), (1, 0.1, alpha), (0, 0.03, alpha))):
        m = [np.full(n, -7.0 - i) for i in range(4)]
        if al is None:
            be.ns_macro_collide(l, *m, nu, issave)
        else:
            be.ns_macro_brinkman_collide(l, *m, nu, al, issave)
        if issave:
            res += [(f"m{k}{i}", m[i].copy()) for i in range(dim + 1)]
        res += S.pops(f"c{k}", l)
    l.free()
    return res


@pytest.mark.parametrize("dim", [2, 3])
def test_scalar_order_site_math_equals_the_scalar_build_of_the_reference(dim):
    import ctypes as C
    ref, hm = _scalar_pair(dim)
    try:
        for n, size in enumerate(SIZES[dim]):
            S.assert_same(_ns_collides(ref, dim, size, 3 + n), _ns_collides(hm, dim, size, 3 + n), f"NS collides {size}")
            for model in S.FORWARD_MODELS + S.ADJOINT_MODELS:
                if model.endswith("massflow") and dim == 3:
                    continue
                S.assert_same(S.collide(ref, dim, model, size, 3 + n), S.collide(hm, dim, model, size, 3 + n), f"{model} {size}")
            for kind in S.CLOSURES:
                if kind == "aad_iset_rho" and dim == 3:
                    continue
                S.assert_same(S.closure(ref, dim, kind, size, 5 + n), S.closure(hm, dim, kind, size, 5 + n), f"{kind} {size}")
            for kind in S.SENSITIVITIES:
                if kind == "aad_temperature_at_heat_source" and dim == 3:
                    continue    # the reference's scalar 3-D overload reads out of bounds (_uz / _ig swapped, adjointadvection.h:1536 vs :805)
                S.assert_same(S.sensitivity(ref, dim, kind, size, 9 + n), S.sensitivity(hm, dim, kind, size, 9 + n), f"{kind} {size}")
            S.assert_same(S.inits(ref, dim, size, 2 + n), S.inits(hm, dim, size, 2 + n), f"init {size}")
    finally:
        hm._fn("set_scalar_build")(C.c_int(0))


def test_planar_sensitivity_on_a_3d_lattice():
    """test/nssens3D.cpp:105 hands its D3Q15 lattice to the two-component overload of ANS::SensitivityBrinkman; pl_sensitivity accepts
    the call (uz / imz absent) and evaluates the two-component expression at every site, as the reference does"""
    if 3 not in DIMS:
        pytest.skip("oracle/_ref (3-D) not built")
    import numpy as np
    from helpers import random_field
    ref, hm = both(3)
    if not ref.has("ans_sensitivity_brinkman_planar"):
        pytest.skip("oracle/_ref predates the entry point: make -C oracle ref")
    for n, size in enumerate(SIZES[3]):
        res = []
        for be in (ref, hm):
            l = be.lattice(*size)
            F = S.Fields(l.nxyz, 9 + n)
            dfds = random_field(l.nxyz, 40 + n, -1, 1)
            be.ans_sensitivity_brinkman_planar(l, dfds, F.ux, F.uy, F.imx, F.imy, F.dads)
            res.append(dfds)
            l.free()
        assert np.array_equal(res[0], res[1]), size
