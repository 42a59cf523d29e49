"""Fixtures for tests/test_gpu_nsin.py: tests/dropin/nsin_dump.cpp (loops over src/equation/nsincompressible.h) built against the
UNMODIFIED reference headers (-I/root/reference/src, README flags + -O2 -ffp-contract=off) and run on the CPU in this container.
    python tests/golden/make_nsin_golden.py        (needs /root/reference; writes tests/golden/nsin.npz)
Per case and output array: SHA-256 of the raw fp64 bytes and every 5th value."""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
# tag -> (lx, ly, nt, dt)
NSIN_CASES = {"nsin": (101, 91, 300, 100), "nsin_small": (43, 37, 120, 50)}


def main():
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    res = {}
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "nsin_ref")
        subprocess.check_call(["g++", "-O2", "-mavx", "-fopenmp", "-ffp-contract=off", "-w", "-I" + os.path.join(REF, "src"),
                               os.path.join(os.path.dirname(HERE), "dropin", "nsin_dump.cpp"), "-o", exe], env=env)
        for tag, args in NSIN_CASES.items():
            w = os.path.join(d, tag)
            os.makedirs(w)
            subprocess.run([exe, *[str(a) for a in args], w], capture_output=True, text=True, check=True)
            for f in sorted(os.listdir(w)):
                if f.endswith(".out"):
                    a = np.fromfile(os.path.join(w, f)) + 0.0
                    res[f"{tag}/{f[:-4]}/sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
                    res[f"{tag}/{f[:-4]}/s5"] = a[::5]
    np.savez_compressed(os.path.join(HERE, "nsin.npz"), **res)
    print(len(res)//2, "arrays", os.path.getsize(os.path.join(HERE, "nsin.npz")), "bytes")


if __name__ == "__main__":
    sys.exit(main())
