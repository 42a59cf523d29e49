"""CPU-side check of the drop-in C++ surface: the parity programs of tests/dropin/ — the heatsink, transient and ncpump loop bodies
written against the reference API — must compile against panslbm2_b200/src and link with libpanslbm_b200.so.  No device call is
made (the GPU suites run them); this catches a signature of the reference API going missing from the headers."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(ROOT, "panslbm2_b200")

PROGRAMS = [("heatsink_dump.cpp", []), ("heatsink_dump.cpp", ["-DHEATSINK_SCALAR"]), ("transient_dump.cpp", ["-DTRANSIENT_DIM=3", "-DPANSLBM_B200_DROPIN", "-I" + os.path.join(LIB, "src")]),
            ("transient_dump.cpp", ["-DTRANSIENT_DIM=2", "-DPANSLBM_B200_DROPIN", "-I" + os.path.join(LIB, "src")]),
            ("ncpump_dump.cpp", ["-DPANSLBM_B200_DROPIN", "-I" + os.path.join(LIB, "src")]), ("filter_dump.cpp", []),
            ("ncpump_periodic_dump.cpp", ["-DPANSLBM_B200_DROPIN", "-I" + os.path.join(LIB, "src")]),
            ("nsin_dump.cpp", ["-DPANSLBM_B200_DROPIN", "-I" + os.path.join(LIB, "src")]),
            ("nsopt_dump.cpp", ["-DPANSLBM_B200_DROPIN", "-I" + os.path.join(LIB, "src")])]


@pytest.mark.parametrize("src,flags", PROGRAMS, ids=[p[0] + "".join(f for f in p[1] if f.startswith(("-DTRANSIENT", "-DHEATSINK"))) for p in PROGRAMS])
def test_parity_program_builds_against_dropin_headers(tmp_path, src, flags):
    if not os.path.exists(os.path.join(LIB, "libpanslbm_b200.so")):
        pytest.skip("libpanslbm_b200.so not built (python __graft_entry__.py)")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    out = str(tmp_path / "prog")
    r = subprocess.run(["g++", "-O1", "-mavx", "-w", *flags, "-I" + os.path.join(ROOT, "include"), os.path.join(HERE, "dropin", src), "-o", out,
                        "-L" + LIB, "-lpanslbm_b200", "-Wl,-rpath," + LIB], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    # without arguments every program prints its usage and exits 2 before touching the device
    u = subprocess.run([out], capture_output=True, text=True)
    assert u.returncode == 2 and "usage" in (u.stderr + u.stdout)
