"""Fixtures for tests/test_gpu_ncpump.py: tests/dropin/ncpump_dump.cpp (the loops of production/ncpump.cpp:112-245) built against the
UNMODIFIED reference headers (-I/root/reference/src, README flags + -O2 -ffp-contract=off) and run on the CPU in this container.
    python tests/golden/make_ncpump_golden.py        (needs /root/reference; writes tests/golden/ncpump.npz)
Per case and output array: SHA-256 of the raw fp64 bytes and every 5th value."""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
# tag -> (lx, ly, nt); 51 x 101 is the driver's own size (ncpump.cpp:41), 5151 sites = 3 in the scalar tail
NCPUMP_CASES = {"ncp": (51, 101, 300), "ncp_small": (31, 42, 150)}


def main():
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    res = {}
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "ncpump_ref")
        subprocess.check_call(["g++", "-O2", "-mavx", "-fopenmp", "-ffp-contract=off", "-w", "-I" + os.path.join(REF, "src"),
                               os.path.join(os.path.dirname(HERE), "dropin", "ncpump_dump.cpp"), "-o", exe], env=env)
        for tag, (lx, ly, nt) in NCPUMP_CASES.items():
            w = os.path.join(d, tag)
            os.makedirs(w)
            r = subprocess.run([exe, str(lx), str(ly), str(nt), w], capture_output=True, text=True, check=True)
            print(tag, r.stdout.strip())
            for f in sorted(os.listdir(w)):
                if f.endswith(".out"):
                    a = np.fromfile(os.path.join(w, f)) + 0.0
                    res[f"{tag}/{f[:-4]}/sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
                    res[f"{tag}/{f[:-4]}/s5"] = a[::5]
    np.savez_compressed(os.path.join(HERE, "ncpump.npz"), **res)
    print(len(res)//2, "arrays")


if __name__ == "__main__":
    sys.exit(main())
