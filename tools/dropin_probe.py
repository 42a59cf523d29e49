"""time tests/dropin/heatsink_dump.cpp (the heatsink loop bodies through the drop-in C++ surface) at a given size:
    python tools/dropin_probe.py <dim> <lx> <ly> <lz> <nt> [ENV=VALUE ...]      prints the program's own timing lines"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import heatsink_case as H
from helpers import gcoords

dim, lx, ly, lz, nt = [int(v) for v in sys.argv[1:6]]
extra = dict(a.split("=", 1) for a in sys.argv[6:])
size = (lx, ly, lz)
env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
lib = os.path.join(ROOT, "panslbm2_b200")
with tempfile.TemporaryDirectory() as d:
    exe = os.path.join(d, "heatsink_dump")
    subprocess.check_call(["g++", "-O2", "-mavx", "-ffp-contract=off", "-w", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "dropin", "heatsink_dump.cpp"),
                           "-o", exe, "-L" + lib, "-lpanslbm_b200", "-Wl,-rpath," + lib], env=env)
    p = H.params(dim, size)
    for name, a in zip(("alpha", "kappa", "dads", "dkds"), H.design_fields(p, *gcoords(*size))):
        np.ascontiguousarray(a, dtype=np.float64).tofile(os.path.join(d, name + ".bin"))
    np.array([p["nu"], p["gx"], p["gy"], p["gz"], p["tem0"], p["qn0"], p["L"]]).tofile(os.path.join(d, "params.bin"))
    e = dict(env); e.update(extra); e["HEATSINK_DUMP_NO_OUTPUT"] = "1"
    r = subprocess.run([exe, str(dim), str(lx), str(ly), str(lz), str(nt), d], capture_output=True, text=True, env=e)
    print(" ".join(sys.argv[1:]), "|", r.stdout.strip().replace("\n", " | "), r.stderr[-300:] if r.returncode else "")
