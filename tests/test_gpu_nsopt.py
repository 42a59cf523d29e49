"""GPU parity of the pipe-bend loops (production/nsopt.cpp:82-150) through the drop-in C++ surface: D2Q9 NS::MacroBrinkmanCollide with a
parabolic SetU inlet, a SetRho outlet and bounce-back elsewhere, the ANS adjoint loop (iBoundaryConditionSetU with eps = 1,
iBoundaryConditionSetRho2D), the pressure-drop objective read from rho, ANS::SensitivityBrinkman and Normalize.
nsopt.cpp is the one reference program that leaves _USE_AVX_DEFINES commented out (:2): built as committed, the reference runs its
scalar templates at every site; with the macro, the AVX overloads.  The two builds of the reference differ from each other by up to
6e-12 (forward fields), 9e-10 (adjoint fields) and 4e-9 (sensitivity) relative L-inf on these cases, and at the last nxyz % 4 sites
in what they store (macros before / after the Brinkman force, navierstokes_avx.h:246-254 vs navierstokes.h:494-503).  The drop-in
headers honour the same macro (pl_set_scalar_order): tests/dropin/nsopt_dump.cpp compiled against panslbm2_b200/src with and
without it must reproduce, bit for bit, the fixtures the same source produced against the reference headers with and without it."""
import hashlib
import importlib.util
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
G = os.path.join(HERE, "golden")


def cases():
    spec = importlib.util.spec_from_file_location("make_nsopt_golden", os.path.join(G, "make_nsopt_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.NSOPT_CASES


@pytest.fixture(scope="session")
def exes(tmp_path_factory):
    d = tmp_path_factory.mktemp("nsopt")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    lib = os.path.join(ROOT, "panslbm2_b200")
    out = {}
    for build, flags in (("avx", ["-DNSOPT_AVX"]), ("scalar", [])):
        out[build] = str(d / ("nsopt_dump_" + build))
        subprocess.check_call(["g++", "-O2", "-mavx", "-ffp-contract=off", "-w", "-DPANSLBM_B200_DROPIN", *flags, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(lib, "src"),
                               os.path.join(HERE, "dropin", "nsopt_dump.cpp"), "-o", out[build], "-L" + lib, "-lpanslbm_b200", "-Wl,-rpath," + lib], env=env)
    return out


@pytest.mark.parametrize("build", ["avx", "scalar"])
@pytest.mark.parametrize("tag", ["pipe", "pipe_small"])
def test_nsopt_loops_match_reference_fixture(exes, tmp_path, tag, build):
    lx, ly, nt, dt = cases()[tag]
    r = subprocess.run([exes[build], str(lx), str(ly), str(nt), str(dt), str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = {f[:-4]: np.fromfile(os.path.join(str(tmp_path), f)) for f in os.listdir(str(tmp_path)) if f.endswith(".out")}
    z = np.load(os.path.join(G, "nsopt.npz"))
    ftag = tag if build == "avx" else tag + ".scalar"
    keys = sorted(k.split("/")[1] for k in z.files if k.startswith(ftag + "/") and k.endswith("/sha"))
    assert len(keys) >= 16
    for k in keys:
        a = res[k] + 0.0
        want = z[f"{ftag}/{k}/s5"]
        if k == "extra":       # [objective read from rho, Residual forward, Residual adjoint]
            assert a.shape == (3,) and a[0] == want[0]
            # the residuals are reductions: the device sums in another order than the host loop (residual.h:8-50)
            assert np.all(np.abs(a[1:] - want[1:]) <= 1e-9*np.abs(want[1:])), (a, want)
            continue
        assert np.array_equal(a[::5], want), f"{ftag}: {k} differs from the reference fixture (max abs {np.max(np.abs(a[::5] - want)):.3e})\n{r.stdout}"
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest() == bytes(z[f"{ftag}/{k}/sha"]), f"{ftag}: {k} digest"
    # both loops replay as fused passes (Residual every dt steps settles the plan and re-learns)
    assert res["stats"][0] >= 2*(nt - 4*(nt//dt) - 8), r.stdout
