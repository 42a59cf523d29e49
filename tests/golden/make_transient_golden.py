"""Fixtures for tests/test_gpu_transient.py (BASELINE configs[4], the transient heatsink loops): tests/dropin/transient_dump.cpp
built against the UNMODIFIED reference headers (-I/root/reference/src) with the reference's flags (README.md:23-25: g++ -mavx
-fopenmp; plus -O2 -ffp-contract=off) and run on the CPU in this container.
    python tests/golden/make_transient_golden.py        (needs /root/reference; writes tests/golden/transient.npz)
Per case and output array: SHA-256 of the raw fp64 bytes and every 5th value."""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import heatsink_case as H          # noqa: E402
from helpers import gcoords        # noqa: E402

REF = "/root/reference"
# tag -> (dim, (lx, ly, lz), nt): nt arrays of every macroscopic field and nt thermal snapshots are stored per case
TRANSIENT_CASES = {"tr3d": (3, (24, 20, 18), 24), "tr3d_tail": (3, (13, 11, 9), 12), "tr2d": (2, (40, 30, 1), 30)}


def write_inputs(d, dim, size):
    p = H.params(dim, size)
    for name, a in zip(("alpha", "kappa", "dads", "dkds"), H.design_fields(p, *gcoords(*size))):
        np.ascontiguousarray(a, dtype=np.float64).tofile(os.path.join(d, name + ".bin"))
    np.array([p["nu"], p["gx"], p["gy"], p["gz"], p["tem0"], p["qn0"], p["L"]]).tofile(os.path.join(d, "params.bin"))


def main():
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for dim in (2, 3):
            subprocess.check_call(["g++", "-O2", "-mavx", "-fopenmp", "-ffp-contract=off", "-w", f"-DTRANSIENT_DIM={dim}", "-I" + os.path.join(REF, "src"),
                                   os.path.join(os.path.dirname(HERE), "dropin", "transient_dump.cpp"), "-o", os.path.join(d, f"transient_ref{dim}")], env=env)
        for tag, (dim, size, nt) in TRANSIENT_CASES.items():
            exe = os.path.join(d, f"transient_ref{dim}")
            w = os.path.join(d, tag)
            os.makedirs(w)
            write_inputs(w, dim, size)
            r = subprocess.run([exe, str(dim), *[str(s) for s in size], str(nt), w], capture_output=True, text=True, check=True)
            print(tag, r.stdout.strip())
            for f in sorted(os.listdir(w)):
                if f.endswith(".out"):
                    a = np.fromfile(os.path.join(w, f)) + 0.0
                    res[f"{tag}/{f[:-4]}/sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
                    res[f"{tag}/{f[:-4]}/s5"] = a[::5]
    np.savez_compressed(os.path.join(HERE, "transient.npz"), **res)
    print(len(res)//2, "arrays")


if __name__ == "__main__":
    sys.exit(main())
