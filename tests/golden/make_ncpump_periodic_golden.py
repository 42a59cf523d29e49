"""Fixtures for tests/test_gpu_ncpump.py::test_ncpump_periodic_*: tests/dropin/ncpump_periodic_dump.cpp (the loops of
production/ncpump_periodic.cpp:107-305) built against the UNMODIFIED reference headers (-I/root/reference/src, README flags + -O2
-ffp-contract=off) and run on the CPU in this container.
    python tests/golden/make_ncpump_periodic_golden.py        (needs /root/reference; writes tests/golden/ncpump_periodic.npz)
Per case and output array: SHA-256 of the raw fp64 bytes and every 5th value (fobj, extra: every value)."""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
# tag -> (lx, ly, nt0, nt, nk); 101 x 201 is the driver's own size (ncpump_periodic.cpp:40), 20 301 sites = 1 in the scalar tail;
# nt0 run-in steps, nt stored steps per period, nk optimisation iterations (the second restarts from the last stored step)
NCPUMP_PERIODIC_CASES = {"ncpp": (101, 201, 300, 120, 2), "ncpp_small": (31, 42, 100, 60, 2)}
WHOLE = ("fobj", "extra")


def main():
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    res = {}
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "ncpump_periodic_ref")
        subprocess.check_call(["g++", "-O2", "-mavx", "-fopenmp", "-ffp-contract=off", "-w", "-I" + os.path.join(REF, "src"),
                               os.path.join(os.path.dirname(HERE), "dropin", "ncpump_periodic_dump.cpp"), "-o", exe], env=env)
        for tag, args in NCPUMP_PERIODIC_CASES.items():
            w = os.path.join(d, tag)
            os.makedirs(w)
            r = subprocess.run([exe, *[str(a) for a in args], w], capture_output=True, text=True, check=True)
            print(tag, r.stdout.strip())
            for f in sorted(os.listdir(w)):
                if f.endswith(".out"):
                    a = np.fromfile(os.path.join(w, f)) + 0.0
                    res[f"{tag}/{f[:-4]}/sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
                    res[f"{tag}/{f[:-4]}/s5"] = a if f[:-4] in WHOLE else a[::5]
    np.savez_compressed(os.path.join(HERE, "ncpump_periodic.npz"), **res)
    print(len(res)//2, "arrays", os.path.getsize(os.path.join(HERE, "ncpump_periodic.npz")), "bytes")


if __name__ == "__main__":
    sys.exit(main())
