"""GPU parity of the transient heatsink loops (BASELINE configs[4]: production/heatsink3D_transient.cpp:145-232,
production/heatsink_transient.cpp:136-215) through the drop-in C++ surface.

tests/dropin/transient_dump.cpp keeps one set of macroscopic arrays and one thermal snapshot PER TIME STEP, walks them backwards
in the adjoint loop with a sensitivity accumulation every step, and sums the objective over tem[t] afterwards — as the drivers do.
Compiled here against panslbm2_b200/src it must reproduce, bit for bit, the fixtures the same source produced against the
reference headers (tests/golden/transient.npz, made by tests/golden/make_transient_golden.py).  The loops must run FUSED (the
plan's array arguments are re-bound every step) and, with a device budget smaller than the stored states, spill the oldest
states to the host and bring them back in the adjoint loop."""
import hashlib
import importlib.util
import os
import subprocess

import numpy as np
import pytest

import heatsink_case as H
from helpers import gcoords

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
G = os.path.join(HERE, "golden")


def cases():
    spec = importlib.util.spec_from_file_location("make_transient_golden", os.path.join(G, "make_transient_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.TRANSIENT_CASES


@pytest.fixture(scope="session")
def exes(tmp_path_factory):
    d = tmp_path_factory.mktemp("transient")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    lib = os.path.join(ROOT, "panslbm2_b200")
    out = {}
    for dim in (2, 3):
        out[dim] = str(d / f"transient_dump{dim}")
        subprocess.check_call(["g++", "-O2", "-mavx", "-ffp-contract=off", "-w", f"-DTRANSIENT_DIM={dim}", "-DPANSLBM_B200_DROPIN", "-I" + os.path.join(ROOT, "include"),
                               "-I" + os.path.join(lib, "src"), os.path.join(HERE, "dropin", "transient_dump.cpp"), "-o", out[dim],
                               "-L" + lib, "-lpanslbm_b200", "-Wl,-rpath," + lib], env=env)
    return out


def run(exe, d, dim, size, nt, env=None):
    p = H.params(dim, size)
    for name, a in zip(("alpha", "kappa", "dads", "dkds"), H.design_fields(p, *gcoords(*size))):
        np.ascontiguousarray(a, dtype=np.float64).tofile(os.path.join(d, name + ".bin"))
    np.array([p["nu"], p["gx"], p["gy"], p["gz"], p["tem0"], p["qn0"], p["L"]]).tofile(os.path.join(d, "params.bin"))
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([exe, str(dim), *[str(s) for s in size], str(nt), d], capture_output=True, text=True, timeout=900, env=e)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return {f[:-4]: np.fromfile(os.path.join(d, f)) for f in os.listdir(d) if f.endswith(".out")}, r.stdout


def check(tag, res):
    z = np.load(os.path.join(G, "transient.npz"))
    keys = sorted(k.split("/")[1] for k in z.files if k.startswith(tag + "/") and k.endswith("/sha"))
    assert len(keys) >= 25
    for k in keys:
        a = res[k] + 0.0
        assert np.array_equal(a[::5], z[f"{tag}/{k}/s5"]), f"{tag}: {k} differs from the reference fixture (max abs {np.max(np.abs(a[::5] - z[f'{tag}/{k}/s5'])):.3e})"
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest() == bytes(z[f"{tag}/{k}/sha"]), f"{tag}: {k} digest"


@pytest.mark.parametrize("tag", ["tr3d", "tr3d_tail", "tr2d"])
def test_transient_loops_match_reference_fixture(exes, tmp_path, tag):
    dim, size, nt = cases()[tag]
    res, log = run(exes[dim], str(tmp_path), dim, size, nt)
    check(tag, res)
    fused = res["stats"][0]
    n = size[0]*size[1]*size[2]
    if n*8 >= 4096:
        # both loops run as fused passes once learned: (nt - 1) forward + (nt - 1) adjoint collides, minus the learning iterations
        assert fused >= 2*(nt - 1) - 8, log


def test_transient_state_store_spills_and_restores(exes, tmp_path):
    """device budget far below the stored states: the oldest states go to the host during the forward loop and come back,
    newest first, in the adjoint loop — same numbers"""
    dim, size, nt = cases()["tr3d"]
    n = size[0]*size[1]*size[2]
    per_step = 23*n*8
    res, log = run(exes[dim], str(tmp_path), dim, size, nt, env={"PANSLBM_B200_DEVICE_BUDGET_MB": str(max(1, 6*per_step >> 20))})
    check("tr3d", res)
    assert "spilled" in log and "spilled 0 " not in log, log


def _ngpu():
    from panslbm2_b200 import _lib
    return _lib.lib().pl_device_count()


@pytest.fixture(scope="session")
def mpi_exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("transient_mpi") / "transient_dump3_mpi")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    lib = os.path.join(ROOT, "panslbm2_b200")
    subprocess.check_call(["g++", "-O2", "-mavx", "-ffp-contract=off", "-w", "-DTRANSIENT_DIM=3", "-DPANSLBM_B200_DROPIN", "-DTRANSIENT_MPI", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(lib, "src"), os.path.join(HERE, "dropin", "transient_dump.cpp"), "-o", out, "-L" + lib, "-lpanslbm_b200", "-Wl,-rpath," + lib], env=env)
    return out


def assemble(d, name, size, nranks):
    """per-rank blocks <name>.r<rank>.out -> the array of the global domain (per-site fields and AoS populations alike)"""
    lx, ly, lz = size
    out = None
    for r in range(nranks):
        ox, oy, oz, nx, ny, nz = [int(v) for v in np.fromfile(os.path.join(d, f"block.r{r}.out"))]
        a = np.fromfile(os.path.join(d, f"{name}.r{r}.out"))
        per = a.size//(nx*ny*nz)
        if out is None:
            out = np.zeros((lz, ly, lx, per))
        out[oz:oz + nz, oy:oy + ny, ox:ox + nx, :] = a.reshape(nz, ny, nx, per)
    return out.reshape(-1)


@pytest.mark.parametrize("pe", ["1,1,2", "2,1,1", "1,2,2", "2,2,2"])
def test_transient_loops_over_ranks_match_reference_fixture(mpi_exe, tmp_path, pe):
    """BASELINE configs[4] on a PE grid (production/heatsink3D_transient.cpp:28-48): one process per GPU under tools/mpiexec_b200,
    halo exchange over NCCL inside the fused, re-bound passes, MPI_Allreduce of the objective through the mpi.h shim.  Blocks of
    the 24 x 20 x 18 case hold multiples of 4 sites: bit-identical to the single-rank reference fixture."""
    n = eval(pe.replace(",", "*"))
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    dim, size, nt = cases()["tr3d"]
    d = str(tmp_path)
    p = H.params(dim, size)
    for name, a in zip(("alpha", "kappa", "dads", "dkds"), H.design_fields(p, *gcoords(*size))):
        np.ascontiguousarray(a, dtype=np.float64).tofile(os.path.join(d, name + ".bin"))
    np.array([p["nu"], p["gx"], p["gy"], p["gz"], p["tem0"], p["qn0"], p["L"]]).tofile(os.path.join(d, "params.bin"))
    env = dict(os.environ, TRANSIENT_PE=pe, PANSLBM_LOG_DIR=d, PANSLBM_RDV_DIR=d)
    r = subprocess.run([os.path.join(ROOT, "tools", "mpiexec_b200"), "-n", str(n), mpi_exe, str(dim), *[str(s) for s in size], str(nt), d],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    z = np.load(os.path.join(G, "transient.npz"))
    keys = sorted(k.split("/")[1] for k in z.files if k.startswith("tr3d/") and k.endswith("/sha"))
    checked = 0
    for k in keys:
        if k in ("stats",):
            continue
        if k == "extra":
            got = np.fromfile(os.path.join(d, "extra.r0.out"))
            assert abs(got[0] - z["tr3d/extra/s5"][0]) <= 1e-12*abs(z["tr3d/extra/s5"][0]), "objective"
            continue
        a = assemble(d, k, size, n) + 0.0
        assert np.array_equal(a[::5], z[f"tr3d/{k}/s5"]), f"{pe}: {k} differs from the reference fixture (max abs {np.max(np.abs(a[::5] - z[f'tr3d/{k}/s5'])):.3e})"
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest() == bytes(z[f"tr3d/{k}/sha"]), f"{pe}: {k} digest"
        checked += 1
    assert checked >= 25
