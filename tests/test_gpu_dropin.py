"""GPU parity of the drop-in C++ surface (panslbm2_b200/src: the reference's header API in front of libpanslbm_b200.so).

1. tests/dropin/heatsink_dump.cpp — the heatsink loop bodies written against the reference API exactly as the drivers do
   (plain new[] arrays, std::swap, direct reads) — is compiled here with g++ and must reproduce, bit for bit, the fixtures the
   reference build generated (tests/golden/heatsink.npz).  That exercises the host runtime end to end: array mirroring with
   page-protection coherence, lambda baking, loop recognition and the fused replay.
2. UNMODIFIED reference programs (test/cavityflow3D.cpp, test/d2q9.cpp, test/d3q15.cpp) built against the drop-in headers by
   tools/build_dropin.sh (binaries under build/dropin/bin, which travel with the snapshot) must print / write what the reference
   build printed / wrote (tests/golden/dropin.npz)."""
import hashlib
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
BIN = os.path.join(ROOT, "build", "dropin", "bin")


@pytest.fixture(scope="session")
def dump_exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dropin") / "heatsink_dump")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    lib = os.path.join(ROOT, "panslbm2_b200")
    subprocess.check_call(["g++", "-O2", "-mavx", "-ffp-contract=off", "-w", "-I" + os.path.join(ROOT, "include"), os.path.join(HERE, "dropin", "heatsink_dump.cpp"),
                           "-o", out, "-L" + lib, "-lpanslbm_b200", "-Wl,-rpath," + lib], env=env)
    return out


def heatsink_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.HEATSINK_CASES


def run_dump(exe, d, dim, size, nt, env=None):
    p = H.params(dim, size)
    for name, a in zip(("alpha", "kappa", "dads", "dkds"), H.design_fields(p, *gcoords(*size))):
        np.ascontiguousarray(a, dtype=np.float64).tofile(os.path.join(d, name + ".bin"))
    np.array([p["nu"], p["gx"], p["gy"], p["gz"], p["tem0"], p["qn0"], p["L"]]).tofile(os.path.join(d, "params.bin"))
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([exe, str(dim), *[str(s) for s in size], str(nt), d], capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return {f[:-4]: np.fromfile(os.path.join(d, f)) for f in os.listdir(d) if f.endswith(".out")}, r.stdout


@pytest.mark.parametrize("knobs", [{"PANSLBM_XGHOST": "0"}, {"PANSLBM_XGHOST": "0", "PANSLBM_XINLINE": "1"}, {"PANSLBM_GRAPH": "1"},
                                   {"PANSLBM_XGHOST": "0", "PANSLBM_SHELL_SERIAL": "1", "PANSLBM_PREFETCH": "3"}, {"PANSLBM_INPLACE": "0"},
                                   {"PANSLBM_INPLACE": "0", "PANSLBM_XGHOST": "0"}, {"PANSLBM_GRAPH": "1", "PANSLBM_XGHOST": "0"}, {"PANSLBM_COOP_SITES": "400000"},
                                   {"PANSLBM_L2_AHEAD": "0"}, {"PANSLBM_PIPE": "1"}],
                         ids=["xslab", "xinline", "graph", "serial_prefetch", "two_buffers", "two_buffers_xslab", "graph_xslab", "coop", "no_l2_ahead", "pipe"])
@pytest.mark.parametrize("tag", ["hs3d", "hs2d"])
def test_alternative_boundary_schedules_give_the_same_numbers(dump_exe, tmp_path, tag, knobs):
    """the x closure planes can be served three ways (k_xclose ahead of the pass on the compact wall buffers = default, aligned x
    groups in the boundary pass, inline in the interior kernel), the step replayed as a CUDA graph, the boundary pass queued behind
    the interior kernel, the passes run from one population buffer into a second one instead of in place: all are schedules of the
    same arithmetic and must reproduce the reference fixture bit for bit"""
    dim, size, nt = heatsink_cases()[tag]
    res, log = run_dump(dump_exe, str(tmp_path), dim, size, nt, env=knobs)
    z = np.load(os.path.join(G, "heatsink.npz"))
    for k in ("rho", "ux", "uy", "tem", "qx", "qy", "ip", "iux", "imx", "item", "iqy", "dfdss", "f.f", "g.f", "f.f0", "g.f0"):
        assert hashlib.sha256(np.ascontiguousarray(res[k] + 0.0).tobytes()).digest() == bytes(z[f"{tag}/{k}/sha"]), f"{tag} {knobs}: {k}"
    assert res["stats"][0] >= 2*(nt - 3), log


@pytest.mark.parametrize("tag", ["hs3d", "hs2d", "hs3d_tail"])
def test_cpp_surface_matches_reference_fixture(dump_exe, tmp_path, tag):
    dim, size, nt = heatsink_cases()[tag]
    res, log = run_dump(dump_exe, str(tmp_path), dim, size, nt)
    z = np.load(os.path.join(G, "heatsink.npz"))
    keys = sorted(k.split("/")[1] for k in z.files if k.startswith(tag + "/") and k.endswith("/sha"))
    checked = 0
    for k in keys:
        # (gsnap / igsnap included: the host sees the thermal snapshots in the reference's own layout, advection_avx.h:1046-1052)
        a = res[k] + 0.0
        assert np.array_equal(a[::5], z[f"{tag}/{k}/s5"]), f"{tag}: {k} differs from the reference fixture (max abs {np.max(np.abs(a[::5] - z[f'{tag}/{k}/s5'])):.3e})"
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest() == bytes(z[f"{tag}/{k}/sha"]), f"{tag}: {k} digest"
        checked += 1
    assert checked >= 19
    # the loops really ran fused: all but the two learning iterations of each loop, and the arrays moved only on demand
    fused, single, uploads, downloads, faults, plans, settles, stagings = res["stats"]
    n = size[0]*size[1]*size[2]
    if n*8 >= 4096:       # PANSLBM_B200_ALLOC_THRESHOLD: smaller arrays stay ordinary heap memory and are staged per call
        assert fused >= 2*(nt - 3), log
        assert plans == 2 and stagings <= 8, log
    # objective read straight from tem[] after the loops (heatsink3D.cpp:227-235)
    L, lx = H.params(dim, size)["L"], size[0]
    tem = res["tem"].reshape(size[2], size[1], size[0])
    want = tem[:int(min(L, size[2])) if dim == 3 else 1, 0, :int(min(L, lx))].sum()
    assert abs(res["extra"][0] - want) <= 1e-14*max(1.0, abs(want))


@pytest.fixture(scope="session")
def dump_exe_scalar(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dropin_scalar") / "heatsink_dump_scalar")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    lib = os.path.join(ROOT, "panslbm2_b200")
    subprocess.check_call(["g++", "-O2", "-mavx", "-ffp-contract=off", "-w", "-DHEATSINK_SCALAR", "-I" + os.path.join(ROOT, "include"), os.path.join(HERE, "dropin", "heatsink_dump.cpp"),
                           "-o", out, "-L" + lib, "-lpanslbm_b200", "-Wl,-rpath," + lib], env=env)
    return out


@pytest.mark.parametrize("tag", ["hs2d", "hs3d_tail"])
def test_cpp_surface_built_without_the_avx_macro_matches_the_scalar_build_of_the_reference(dump_exe_scalar, tmp_path, tag):
    """a program that leaves _USE_AVX_DEFINES out (production/nsopt.cpp:2) gets the arithmetic of the reference's scalar templates at
    every site (pl_set_scalar_order, called by the headers): the heatsink iteration — two lattices, SetT / SetQ closures, thermal
    snapshots in the scalar build's [idx][c] layout, sensitivity — against the reference headers compiled without the macro
    (tests/golden/heatsink_scalar.npz), bit for bit"""
    dim, size, nt = heatsink_cases()[tag]
    res, log = run_dump(dump_exe_scalar, str(tmp_path), dim, size, nt)
    z = np.load(os.path.join(G, "heatsink_scalar.npz"))
    keys = sorted(k.split("/")[1] for k in z.files if k.startswith(tag + "/") and k.endswith("/sha"))
    assert len(keys) >= 18
    for k in keys:
        a = res[k] + 0.0
        assert np.array_equal(a[::5], z[f"{tag}/{k}/s5"]), f"{tag} (scalar order): {k} differs from the reference fixture (max abs {np.max(np.abs(a[::5] - z[f'{tag}/{k}/s5'])):.3e})"
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest() == bytes(z[f"{tag}/{k}/sha"]), f"{tag} (scalar order): {k} digest"
    # ... and it differs from the AVX build's fixture in the last bits, as the reference's own two builds do
    za = np.load(os.path.join(G, "heatsink.npz"))
    assert not np.array_equal(res["ux"][::5], za[f"{tag}/ux/s5"])
    n = size[0]*size[1]*size[2]
    if n*8 >= 4096:
        assert res["stats"][0] >= 2*(nt - 3), log


def test_cpp_surface_production_size_matches_reference_fixture(dump_exe, tmp_path):
    """the same program at 81 x 161 x 81 (production/heatsink3D.cpp:42), 2000 + 2000 steps: digests of the reference build"""
    z = np.load(os.path.join(G, "heatsink_fullsize.npz"))
    lx, ly, lz, nt = [int(v) for v in z["shape"]]
    res, log = run_dump(dump_exe, str(tmp_path), 3, (lx, ly, lz), nt)
    for k in sorted(k[:-4] for k in z.files if k.endswith("/sha")):
        a = res[k] + 0.0
        assert np.array_equal(a[::997], z[f"{k}/s997"]), k
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest() == bytes(z[f"{k}/sha"]), k
    assert res["stats"][0] >= 2*(nt - 3), log


def test_cpp_surface_2d_production_size_matches_reference_fixture(dump_exe, tmp_path):
    """BASELINE configs[1]: the 2-D heatsink iteration at 141 x 161 (production/heatsink.cpp:41), 2000 + 2000 steps: digests of the
    reference build"""
    z = np.load(os.path.join(G, "heatsink2d_fullsize.npz"))
    lx, ly, lz, nt = [int(v) for v in z["shape"]]
    res, log = run_dump(dump_exe, str(tmp_path), 2, (lx, ly, lz), nt)
    for k in sorted(k[:-4] for k in z.files if k.endswith("/sha")):
        a = res[k] + 0.0
        assert np.array_equal(a[::997], z[f"{k}/s997"]), k
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest() == bytes(z[f"{k}/sha"]), k
    assert res["stats"][0] >= 2*(nt - 3), log


def test_cpp_surface_without_alloc_hook_still_correct(tmp_path):
    """PANSLBM_B200_NO_ALLOC_HOOK: arrays are foreign memory, staged around every call — slow path, same numbers"""
    out = str(tmp_path / "heatsink_dump_nohook")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    lib = os.path.join(ROOT, "panslbm2_b200")
    subprocess.check_call(["g++", "-O2", "-mavx", "-ffp-contract=off", "-w", "-DPANSLBM_B200_NO_ALLOC_HOOK", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(HERE, "dropin", "heatsink_dump.cpp"), "-o", out, "-L" + lib, "-lpanslbm_b200", "-Wl,-rpath," + lib], env=env)
    dim, size, nt = heatsink_cases()["hs3d_tail"]
    res, _ = run_dump(out, str(tmp_path), dim, size, nt)
    z = np.load(os.path.join(G, "heatsink.npz"))
    for k in ("rho", "ux", "tem", "qy", "ip", "imx", "item", "iqy", "dfdss", "f.f", "g.f"):
        assert hashlib.sha256(np.ascontiguousarray(res[k] + 0.0).tobytes()).digest() == bytes(z[f"hs3d_tail/{k}/sha"]), k


# ---------------------------------------------------------------------------------------------------------
def need(prog):
    exe = os.path.join(BIN, prog)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} absent: run tools/build_dropin.sh where the reference tree is available")
    return exe


@pytest.mark.parametrize("prog", ["d2q9", "d3q15"])
def test_unmodified_reference_layout_tests(prog):
    """test/d2q9.cpp, test/d3q15.cpp: write f0/f, LoadF, StoreF, print — the reference's own known-answer tests"""
    r = subprocess.run([need(prog)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    want = bytes(np.load(os.path.join(G, "dropin.npz"))[prog + ".stdout"]).decode()
    assert r.stdout == want


def test_unmodified_reference_cavityflow3d(tmp_path):
    """test/cavityflow3D.cpp (31^3, 1000 steps) unmodified: its VTK point data equals the reference build's to the 6 digits written"""
    import re
    os.makedirs(tmp_path / "result")
    r = subprocess.run([need("cavityflow3D")], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    txt = open(tmp_path / "result" / "cavity3D_0.vts").read()
    z = np.load(os.path.join(G, "dropin.npz"))
    for m in re.finditer(r'<DataArray type="Float64" Name="(\w+)" NumberOfComponents="(\d)" format="ascii">(.*?)</DataArray>', txt, re.S):
        got = np.array(m.group(3).split(), dtype=np.float64).reshape(-1, int(m.group(2)))
        want = z["cavity3D." + m.group(1)]
        # ostream's 6 significant digits: identical doubles print identically; allow nothing else
        assert np.array_equal(got, want), (m.group(1), float(np.max(np.abs(got - want))))


# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="session")
def filter_exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dropin_f") / "filter_dump")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    lib = os.path.join(ROOT, "panslbm2_b200")
    subprocess.check_call(["g++", "-O2", "-mavx", "-ffp-contract=off", "-w", "-I" + os.path.join(ROOT, "include"), os.path.join(HERE, "dropin", "filter_dump.cpp"),
                           "-o", out, "-L" + lib, "-lpanslbm_b200", "-Wl,-rpath," + lib], env=env)
    return out


@pytest.mark.parametrize("tag", ["hs3d_box", "hs2d_box", "cone3d", "cone2d_r3"])
def test_gpu_filters_match_reference_filters(filter_exe, tmp_path, tag):
    """DensityFilter / HeavisideFilter through the drop-in headers (GPU kernels) vs the reference's serial host loops.
    The weighted sums are bit-identical (same order, no FMA); the tanh projection differs by the libm/CUDA rounding of tanh."""
    import filter_case as FC
    dim, size, R, beta, box = FC.CASES[tag]
    v, d = FC.inputs(tag)
    v.tofile(tmp_path / "v.bin"); d.tofile(tmp_path / "d.bin")
    b = box or (0, 0, 0)
    r = subprocess.run([filter_exe, str(dim), *[str(s) for s in size], repr(R), repr(beta), *[str(x) for x in b], str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    z = np.load(os.path.join(G, "filters.npz"))
    fv, rho, dfds = [np.fromfile(tmp_path / (n + ".out")) for n in ("fv", "rho", "dfds")]
    assert np.array_equal(fv, z[f"{tag}/fv"])                                  # no transcendental: bit-exact
    assert np.max(np.abs(rho - z[f"{tag}/rho"])) <= 1e-14
    assert np.max(np.abs(dfds - z[f"{tag}/dfds"])) <= 1e-13*np.max(np.abs(z[f"{tag}/dfds"]))


def test_unmodified_reference_nsadncsens(tmp_path):
    """test/nsadncsens.cpp unmodified (D2Q9 71 x 81, fixed two-block design, 100 000 forward + 100 000 adjoint steps through the
    learned/fused replay, SensitivityTemperatureAtHeatSource, Normalize, the reference's VTK writer): every array it writes equals
    the reference build's output to the 6 digits written."""
    import re
    os.makedirs(tmp_path / "result")
    r = subprocess.run([need("nsadncsens")], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    txt = open(tmp_path / "result" / "nsadncsens_0.vts").read()
    z = np.load(os.path.join(G, "dropin.npz"))
    seen = 0
    for m in re.finditer(r'<DataArray type="Float64" Name="(\w+)" NumberOfComponents="(\d)" format="ascii">(.*?)</DataArray>', txt, re.S):
        got = np.array(m.group(3).split(), dtype=np.float64).reshape(-1, int(m.group(2)))
        want = z["nsadncsens." + m.group(1)]
        assert np.array_equal(got, want), (m.group(1), float(np.max(np.abs(got - want))), float(np.max(np.abs(want))))
        seen += 1
    assert seen == len([k for k in z.files if k.startswith("nsadncsens.")]) and seen >= 10


# nssens and nsadsens passed on a B200 in the last GPU window of round 2.  nssens3D aborted there: it calls the 2-D overload of
# ANS::SensitivityBrinkman on its D3Q15 lattice (test/nssens3D.cpp:105), which pl_sensitivity refused; it accepts that call now, but the
# GPU budget of the round was spent before the program could be run again — expected to pass, not yet seen to (non-strict xfail).
# test/naturalconvection.cpp (100 000 steps; fixture in dropin_more.npz) did not finish inside that window and is not run here.
@pytest.mark.parametrize("prog,stem", [("nssens", "adjoint"), ("nsadsens", "nsadsens"),
                                       pytest.param("nssens3D", "adjoint3D", marks=pytest.mark.xfail(strict=False, reason="fix not yet run on a GPU (round-2 budget spent)"))])
def test_unmodified_reference_test_programs(tmp_path, prog, stem):
    """test/nssens.cpp (NS + ANS on a 101 x 51 channel, closures capturing by reference), test/nsadsens.cpp (the heat-exchange
    collides and AAD::SensitivityHeatExchange, 30 000 + 30 000 steps) and test/nssens3D.cpp (D3Q15 101 x 51 x 51) unmodified, through
    the learned / fused replay: every array their VTK writer puts out equals the reference build's to the 6 digits written
    (tests/golden/make_dropin_more_golden.py)."""
    import re
    fixture = os.path.join(G, "dropin_more.npz")
    if not os.path.exists(fixture):
        pytest.skip("tests/golden/dropin_more.npz absent")
    z = np.load(fixture)
    os.makedirs(tmp_path / "result")
    r = subprocess.run([need(prog)], cwd=tmp_path, capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stderr[-2000:]
    txt = open(tmp_path / "result" / (stem + "_0.vts")).read()
    seen = 0
    for m in re.finditer(r'<DataArray type="Float64" Name="(\w+)" NumberOfComponents="(\d)" format="ascii">(.*?)</DataArray>', txt, re.S):
        got = np.array(m.group(3).split(), dtype=np.float64).reshape(-1, int(m.group(2)))
        key = f"{prog}.{m.group(1)}"
        if key in z.files:
            want = z[key]
            assert np.array_equal(got, want), (prog, m.group(1), float(np.max(np.abs(got - want))), float(np.max(np.abs(want))))
        else:       # large arrays: every 11th site and the digest of the whole array
            assert np.array_equal(got[::11], z[key + "/s11"]), (prog, m.group(1), float(np.max(np.abs(got[::11] - z[key + "/s11"]))))
            assert hashlib.sha256(np.ascontiguousarray(got + 0.0).tobytes()).digest() == bytes(z[key + "/sha"]), (prog, m.group(1))
        seen += 1
    assert seen == len({k.split("/")[0] for k in z.files if k.startswith(prog + ".")}) and seen >= 2


# ---------------------------------------------------------------------------------------------------------
def vts_pieces(result_dir, stem, names):
    """assemble the per-rank .vts pieces of a VTKXMLExport run into global arrays (pieces overlap by one layer)"""
    import glob
    import re
    pv = open(os.path.join(result_dir, stem + ".pvts")).read()
    ext = [int(x) for x in re.search(r'WholeExtent="([-\d ]+)"', pv).group(1).split()]
    shape = (ext[5] - ext[4] + 1, ext[3] - ext[2] + 1, ext[1] - ext[0] + 1)
    out = {n: np.full(shape, np.nan) for n in names}
    for f in sorted(glob.glob(os.path.join(result_dir, stem + "_*.vts"))):
        txt = open(f).read()
        e = [int(x) for x in re.search(r'<Piece Extent="([-\d ]+)"', txt).group(1).split()]
        ps = (e[5] - e[4] + 1, e[3] - e[2] + 1, e[1] - e[0] + 1)
        for m in re.finditer(r'<DataArray type="Float64" Name="(\w+)" NumberOfComponents="1" format="ascii">(.*?)</DataArray>', txt, re.S):
            if m.group(1) in out:
                a = np.array(m.group(2).split(), dtype=np.float64).reshape(ps)
                out[m.group(1)][e[4]:e[5] + 1, e[2]:e[3] + 1, e[0]:e[1] + 1] = a
    return out


@pytest.mark.parametrize("grid", ["1 1 1", "2 1 1", "1 2 2", "2 2 2"])
def test_unmodified_mpi_program_heavisidefilter(tmp_path, grid):
    """test/heavisidefilter.cpp (hard-wired _USE_MPI_DEFINES; 51^3, R = 1.8) unmodified, over panslbm2_b200/src/mpi/mpi.h: one
    process per GPU, filter on the decomposed lattice, the reference's own VTK writer exchanging through the shim."""
    from panslbm2_b200 import _lib
    n = eval(grid.replace(" ", "*"))
    if _lib.lib().pl_device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    exe = need("heavisidefilter")
    os.makedirs(tmp_path / "result")
    env = dict(os.environ, PANSLBM_RDV_DIR=str(tmp_path), MASTER_PORT=str(29600 + n))
    r = subprocess.run([os.path.join(ROOT, "tools", "mpiexec_b200"), "-n", str(n), exe, *grid.split()], cwd=tmp_path, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = vts_pieces(str(tmp_path / "result"), "heavisidefilter", ("v", "fv"))
    z = np.load(os.path.join(G, "dropin.npz"))
    for k in ("v", "fv"):
        want = z["heavisidefilter." + k].reshape(got[k].shape)
        assert not np.isnan(got[k]).any()
        assert np.max(np.abs(got[k] - want)) <= 2e-6*np.max(np.abs(want)), k       # 6 digits written; tanh rounds differently on the GPU
