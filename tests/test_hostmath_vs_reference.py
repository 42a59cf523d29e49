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
