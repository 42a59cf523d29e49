"""Closures whose values change while the loop runs, through the drop-in C++ surface (the baked-plane cache was keyed by
the raw bytes of the closure objects only).  tests/dropin/closure_dump.cpp drives a D2Q9 channel whose inlet velocity changes every
step — captured by reference, captured by value in a lambda re-created every step, read through a captured pointer — and must
reproduce the fixtures the same source produced with the reference headers (tests/golden/closure.npz), bit for bit, while the loop
still runs fused (a volatile call site keeps its plane handle; the values are replaced in place)."""
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
    spec = importlib.util.spec_from_file_location("make_closure_golden", os.path.join(G, "make_closure_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CASES


@pytest.fixture(scope="session")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("closure") / "closure_dump")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    lib = os.path.join(ROOT, "panslbm2_b200")
    subprocess.check_call(["g++", "-O2", "-mavx", "-ffp-contract=off", "-w", "-DPANSLBM_B200_DROPIN", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(lib, "src"),
                           os.path.join(HERE, "dropin", "closure_dump.cpp"), "-o", out, "-L" + lib, "-lpanslbm_b200", "-Wl,-rpath," + lib], env=env)
    return out


@pytest.mark.parametrize("tag,env", [("byref", {}), ("byvalue", {}), ("pointer", {"PANSLBM_B200_REVALIDATE": "1"})])
def test_changing_closure_values_match_reference_fixture(exe, tmp_path, tag, env):
    mode, lx, ly, nt = cases()[tag]
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([exe, str(mode), str(lx), str(ly), str(nt), str(tmp_path)], capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    z = np.load(os.path.join(G, "closure.npz"))
    for k in ("rho", "ux", "uy", "f.f0", "f.f"):
        got = np.fromfile(os.path.join(str(tmp_path), k + ".out"))
        assert np.array_equal(got, z[f"{tag}/{k}"]), f"{tag}: {k} differs from the reference fixture (max abs {np.max(np.abs(got - z[f'{tag}/{k}'])):.3e})"
    stats = np.fromfile(os.path.join(str(tmp_path), "stats.out"))
    # the loop ran fused although the inlet plane changed every step (its handle survives, the values are replaced in place)
    assert stats[0] >= nt - 30, stats
