"""Fixtures for tests/test_gpu_closure_values.py: tests/dropin/closure_dump.cpp built against the UNMODIFIED reference headers
(-I/root/reference/src, README flags -mavx -fopenmp plus -O2 -ffp-contract=off) and run on the CPU in this container.
    python tests/golden/make_closure_golden.py        (needs /root/reference; writes tests/golden/closure.npz)"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
CASES = {"byref": (0, 44, 26, 150), "byvalue": (1, 44, 26, 150), "pointer": (2, 44, 26, 150)}


def main():
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    res = {}
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "closure_ref")
        subprocess.check_call(["g++", "-O2", "-mavx", "-fopenmp", "-ffp-contract=off", "-w", "-I" + os.path.join(REF, "src"),
                               os.path.join(os.path.dirname(HERE), "dropin", "closure_dump.cpp"), "-o", exe], env=env)
        for tag, (mode, lx, ly, nt) in CASES.items():
            w = os.path.join(d, tag)
            os.makedirs(w)
            subprocess.run([exe, str(mode), str(lx), str(ly), str(nt), w], check=True)
            for f in sorted(os.listdir(w)):
                if f.endswith(".out"):
                    res[f"{tag}/{f[:-4]}"] = np.fromfile(os.path.join(w, f))
    np.savez_compressed(os.path.join(HERE, "closure.npz"), **res)
    print(sorted(res), "max ux", max(float(np.abs(v).max()) for k, v in res.items() if k.endswith("/ux")))


if __name__ == "__main__":
    sys.exit(main())
