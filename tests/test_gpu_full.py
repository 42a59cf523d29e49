"""GPU parity (through the C-ABI) for the thermal / adjoint half of the path: every Macro*Collide* model, every boundary
closure, InitialCondition and sensitivity of NS/AD/ANS/AAD, and the heatsink forward+adjoint+sensitivity iteration
call-by-call and through the fused plan.  Checkers: the committed fixtures generated from the reference build
(tests/golden) and, where oracle/_ref travelled to this box, the reference build itself.  Bit-exact."""
import hashlib
import os

import numpy as np
import pytest

import heatsink_case as H
import scenarios as S
from oracle import oracle as O
from test_golden_ops import DIGESTS, digest, run

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SIZES = {2: [(8, 6, 1), (9, 6, 1), (37, 19, 1)], 3: [(6, 4, 4), (7, 3, 3), (19, 11, 7)]}


def cuda(dim):
    from cuda_ops import CudaOps
    return CudaOps(dim)


@pytest.mark.parametrize("key", sorted(DIGESTS))
def test_ops_match_reference_fixture(key):
    assert digest(run(cuda(int(key[1])), key)) == DIGESTS[key], key


def _vs_ref(dim, fn):
    if not O.have_ref(dim):
        pytest.skip("oracle/_ref did not travel to this box; the fixtures carry the pin")
    ref, cu = O.Backend("ref", dim), cuda(dim)
    for n, size in enumerate(SIZES[dim]):
        S.assert_same(fn(ref, size, n), fn(cu, size, n), str(size))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("model", S.FORWARD_MODELS + S.ADJOINT_MODELS)
def test_collide_models_vs_reference(dim, model):
    if model.endswith("massflow") and dim == 3:
        pytest.skip("the reference's D3Q15 MassFlow overload does not compile")
    _vs_ref(dim, lambda be, size, n: S.collide(be, dim, model, size, 30 + n))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("kind", S.CLOSURES)
def test_closures_vs_reference(dim, kind):
    if kind == "aad_iset_rho" and dim == 3:
        pytest.skip("the reference's D3Q15 AAD::iBoundaryConditionSetRho does not compile")
    _vs_ref(dim, lambda be, size, n: S.closure(be, dim, kind, size, 50 + n))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("kind", S.SENSITIVITIES)
def test_sensitivities_vs_reference(dim, kind):
    _vs_ref(dim, lambda be, size, n: S.sensitivity(be, dim, kind, size, 70 + n))


@pytest.mark.parametrize("dim", [2, 3])
def test_inits_vs_reference(dim):
    _vs_ref(dim, lambda be, size, n: S.inits(be, dim, size, 90 + n))


# the same against the plain-C restatement (oracle/lbm_oracle.c), which always travels with the repository
def _vs_oracle(dim, fn):
    orc, cu = O.Backend("orc", dim), cuda(dim)
    for n, size in enumerate(SIZES[dim]):
        S.assert_same(fn(orc, size, n), fn(cu, size, n), str(size))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("model", S.FORWARD_MODELS + S.ADJOINT_MODELS)
def test_collide_models_vs_c_oracle(dim, model):
    if model.endswith("massflow") and dim == 3:
        pytest.skip("the reference's D3Q15 MassFlow overload does not compile")
    _vs_oracle(dim, lambda be, size, n: S.collide(be, dim, model, size, 130 + n))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("kind", S.CLOSURES)
def test_closures_vs_c_oracle(dim, kind):
    if kind == "aad_iset_rho" and dim == 3:
        pytest.skip("the reference's D3Q15 AAD::iBoundaryConditionSetRho does not compile")
    _vs_oracle(dim, lambda be, size, n: S.closure(be, dim, kind, size, 150 + n))


@pytest.mark.parametrize("dim", [2, 3])
def test_inits_and_sensitivities_vs_c_oracle(dim):
    _vs_oracle(dim, lambda be, size, n: S.inits(be, dim, size, 190 + n))
    for kind in S.SENSITIVITIES:
        _vs_oracle(dim, lambda be, size, n: S.sensitivity(be, dim, kind, size, 170 + n))


@pytest.mark.parametrize("dim,size,nt", [(3, (11, 10, 9), 25), (2, (18, 15, 1), 31)])
def test_heatsink_fused_equals_c_oracle(dim, size, nt):
    H.compare(H.run_cuda(dim, size, nt, fused=True, chunks=(2, 5)), H.run_oplevel(O.Backend("orc", dim), dim, size, nt), "cuda fused vs C oracle")


# ---------------------------------------------------------------------------------------------------------
def check_fixture(tag, res):
    z = np.load(os.path.join(G, "heatsink.npz"))
    keys = sorted(k.split("/")[1] for k in z.files if k.startswith(tag + "/") and k.endswith("/sha"))
    assert keys == sorted(res), (keys, sorted(res))
    for k in keys:
        a = res[k] + 0.0
        assert np.array_equal(a[::5], z[f"{tag}/{k}/s5"]), f"{tag}: {k} differs from the reference fixture (max abs {np.max(np.abs(a[::5] - z[f'{tag}/{k}/s5'])):.3e})"
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest() == bytes(z[f"{tag}/{k}/sha"]), f"{tag}: {k} digest"


def heatsink_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.HEATSINK_CASES


@pytest.mark.parametrize("tag", ["hs3d", "hs2d", "hs3d_tail"])
@pytest.mark.parametrize("fused", [False, True])
def test_heatsink_iteration_matches_reference_fixture(tag, fused):
    dim, size, nt = heatsink_cases()[tag]
    check_fixture(tag, H.run_cuda(dim, size, nt, fused))


@pytest.mark.parametrize("dim,size,nt", [(3, (11, 10, 9), 25), (2, (18, 15, 1), 31)])
def test_heatsink_fused_equals_stepwise_equals_reference(dim, size, nt):
    a = H.run_cuda(dim, size, nt, fused=False)
    b = H.run_cuda(dim, size, nt, fused=True, chunks=(2, 5))
    H.compare(b, a, "fused vs stepwise")
    if O.have_ref(dim):
        H.compare(a, H.run_oplevel(O.Backend("ref", dim), dim, size, nt), "cuda vs reference")


def check_fullsize(res):
    z = np.load(os.path.join(G, "heatsink_fullsize.npz"))
    keys = sorted(k[:-4] for k in z.files if k.endswith("/sha"))
    for k in keys:
        a = res[k] + 0.0
        assert np.array_equal(a[::997], z[f"{k}/s997"]), f"full size: {k} differs from the reference fixture (max abs {np.max(np.abs(a[::997] - z[f'{k}/s997'])):.3e})"
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest() == bytes(z[f"{k}/sha"]), f"full size: {k} digest"
    return len(keys)


def test_heatsink3d_production_size_matches_reference_fixture():
    """81 x 161 x 81 (production/heatsink3D.cpp:42), 2000 forward + 2000 adjoint fused steps + sensitivity: every field bit-identical
    to the reference build's (tests/golden/make_fullsize_golden.py)."""
    z = np.load(os.path.join(G, "heatsink_fullsize.npz"))
    lx, ly, lz, nt = [int(v) for v in z["shape"]]
    assert check_fullsize(H.run_cuda(3, (lx, ly, lz), nt, fused=True, chunks=(100, 100))) >= 24
