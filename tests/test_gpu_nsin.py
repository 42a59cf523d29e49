"""GPU parity of the incompressible NS model (src/equation/nsincompressible.h; SURVEY §8 f4 "remaining collide variants"): collide
models PL_NSIN_COLLIDE / PL_NSIN_BRINKMAN, closures PL_BC_NSIN_SET_U / PL_BC_NSIN_SET_RHO, InitialCondition family 5 — D2Q9 only, as in
the reference.
  * op level through the C-ABI against the C oracle (itself pinned to the reference headers by tests/test_oracle_vs_reference.py);
  * a fused plan against the same calls issued one by one;
  * through the drop-in C++ headers: tests/dropin/nsin_dump.cpp against the fixtures the same source produced with the reference headers."""
import hashlib
import importlib.util
import os
import subprocess

import numpy as np
import pytest

import scenarios as S
from cuda_ops import CudaOps
from helpers import gcoords, same
from oracle import oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
G = os.path.join(HERE, "golden")


@pytest.mark.parametrize("size", [(8, 6, 1), (7, 5, 1), (23, 17, 1), (64, 33, 1)])
def test_nsin_ops_vs_c_oracle(size):
    S.assert_same(S.nsin(O.Backend("orc", 2), size, 4), S.nsin(CudaOps(2), size, 4), f"NSin {size}")


def test_nsin_is_d2q9_only():
    import panslbm2_b200 as pl
    from panslbm2_b200 import api
    p = pl.D3Q15(6, 5, 4)
    a = pl.DeviceArray(p.nxyz, 1.0)
    with pytest.raises(pl.PanslbmError):
        api._collide(p, None, pl.collide_args(api.M_NSIN_COLLIDE, True, 0.1, rho=a, ux=a, uy=a, uz=a))


def nsin_loop(lx, ly, nt, fused):
    """NSin::MacroBrinkmanCollide - Stream - bounce - SetU (parabolic inlet on xmin) - SetRho (outlet on ymin) - SmoothCorner"""
    import panslbm2_b200 as pl
    from panslbm2_b200 import api
    pf = pl.D2Q9(lx, ly)
    N = pf.nxyz
    i, j, _ = gcoords(lx, ly, 1)
    alpha = pl.DeviceArray.from_host(0.02*(1.0 + np.sin(0.3*i)*np.cos(0.2*j)))
    rho, ux, uy = pl.DeviceArray(N, 1.0), pl.DeviceArray(N, 0.0), pl.DeviceArray(N, 0.0)
    inlet = lambda i, j: (i == 0) & (0.3*ly < j) & (j < 0.7*ly)
    outlet = lambda i, j: (j == 0) & (0.5*lx < i) & (i < 0.8*lx)
    wall = lambda i, j: np.where(inlet(i, j) | outlet(i, j), 0, 1)
    uin = [lambda i, j: -0.02*(j - 0.3*ly)*(j - 0.7*ly)/(0.2*ly*0.2*ly), lambda i, j: 0.0*i]
    rout = [lambda i, j: 1.0 + 0.0*i, lambda i, j: 0.0*i]
    pl.NSin.InitialCondition(pf, rho, ux, uy)
    if not fused:
        for _ in range(nt):
            pl.NSin.MacroBrinkmanCollide(pf, rho, ux, uy, 0.1, alpha, True)
            pf.Stream()
            pf.BoundaryCondition(wall)
            pl.NSin.BoundaryConditionSetU(pf, *uin, inlet)
            pl.NSin.BoundaryConditionSetRho(pf, *rout, outlet)
            pf.SmoothCorner()
    else:
        plan = pl.StepPlan(pf)
        plan.set_collide(pl.collide_args(api.M_NSIN_BRINKMAN, True, 0.1, rho=rho, ux=ux, uy=uy, alpha=alpha))
        plan.add_bounce(pf, wall)
        plan.add_closure(pf, api.BC_NSIN_SET_U, inlet, uin)
        plan.add_closure(pf, api.BC_NSIN_SET_RHO, outlet, rout)
        plan.set_smooth_corner(True).finalize()
        first = nt//3
        plan.advance(first, end_streamed=False)
        plan.advance(nt - first, end_streamed=True)
    return [rho.to_host(), ux.to_host(), uy.to_host()], pf.get_populations()


@pytest.mark.parametrize("shape", [(24, 19, 40), (65, 47, 90)])
def test_nsin_fused_equals_stepwise(shape):
    lx, ly, nt = shape
    a, pa = nsin_loop(lx, ly, nt, fused=False)
    b, pb = nsin_loop(lx, ly, nt, fused=True)
    for x, y in zip(a, b):
        assert same(x, y)
    assert same(pa[0], pb[0]) and same(pa[1], pb[1])
    assert np.max(np.abs(a[1])) > 1e-3


def cases():
    spec = importlib.util.spec_from_file_location("make_nsin_golden", os.path.join(G, "make_nsin_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.NSIN_CASES


@pytest.fixture(scope="session")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("nsin") / "nsin_dump")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    lib = os.path.join(ROOT, "panslbm2_b200")
    subprocess.check_call(["g++", "-O2", "-mavx", "-ffp-contract=off", "-w", "-DPANSLBM_B200_DROPIN", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(lib, "src"),
                           os.path.join(HERE, "dropin", "nsin_dump.cpp"), "-o", out, "-L" + lib, "-lpanslbm_b200", "-Wl,-rpath," + lib], env=env)
    return out


@pytest.mark.parametrize("tag", ["nsin_small", "nsin"])
def test_nsin_dropin_loops_match_reference_fixture(exe, tmp_path, tag):
    lx, ly, nt, dt = cases()[tag]
    r = subprocess.run([exe, str(lx), str(ly), str(nt), str(dt), str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = {f[:-4]: np.fromfile(os.path.join(str(tmp_path), f)) for f in os.listdir(str(tmp_path)) if f.endswith(".out")}
    z = np.load(os.path.join(G, "nsin.npz"))
    keys = sorted(k.split("/")[1] for k in z.files if k.startswith(tag + "/") and k.endswith("/sha"))
    assert len(keys) >= 13
    for k in keys:
        a = res[k] + 0.0
        if k == "extra":       # Residual: a reduction, summed in another order than the host loop
            want = z[f"{tag}/extra/s5"]
            assert abs(a[0] - want[0]) <= 1e-9*abs(want[0])
            continue
        assert np.array_equal(a[::5], z[f"{tag}/{k}/s5"]), f"{tag}: {k} differs from the reference fixture (max abs {np.max(np.abs(a[::5] - z[f'{tag}/{k}/s5'])):.3e})\n{r.stdout}"
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest() == bytes(z[f"{tag}/{k}/sha"]), f"{tag}: {k} digest"
    # both loops replay as fused passes
    assert res["stats"][0] >= (nt - 4*(nt//dt) - 8) + (nt//2 - 4), r.stdout
