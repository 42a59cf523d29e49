"""More fixtures for tests/test_gpu_dropin.py: the point data written by four further UNMODIFIED reference programs built with the
reference's own headers and flags (README.md:23-25: g++ -mavx -fopenmp; plus -O2 -ffp-contract=off) and run on the CPU in this
container -> tests/golden/dropin_more.npz (6 significant digits, as the reference's VTK writer prints them).
    python tests/golden/make_dropin_more_golden.py          (needs /root/reference; a few minutes)
test/nssens.cpp (D2Q9 101 x 51 channel with a block, 10 000 + 10 000 steps, ANS adjoint, [&] closures), test/nssens3D.cpp (D3Q15
101 x 51 x 51), test/nsadsens.cpp (heat-exchange collides, 30 000 + 30 000 steps), test/naturalconvection.cpp (100 000 steps)."""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
PROGRAMS = {"nssens": "adjoint", "nssens3D": "adjoint3D", "nsadsens": "nsadsens", "naturalconvection": "naturalconvection"}     # program -> VTK stem


def vts_arrays(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r'<DataArray type="Float64" Name="(\w+)" NumberOfComponents="(\d)" format="ascii">(.*?)</DataArray>', txt, re.S):
        out[m.group(1)] = np.array(m.group(3).split(), dtype=np.float64).reshape(-1, int(m.group(2)))
    return out


def main():
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    res = {}
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "result"))
        for prog, stem in PROGRAMS.items():
            exe = os.path.join(d, prog)
            subprocess.check_call(["g++", "-O2", "-mavx", "-fopenmp", "-ffp-contract=off", "-w", os.path.join(REF, "test", prog + ".cpp"), "-o", exe], env=env)
            subprocess.run([exe], cwd=d, capture_output=True, text=True, check=True)
            arrs = vts_arrays(os.path.join(d, "result", stem + "_0.vts"))
            assert arrs, prog
            for k, v in arrs.items():
                if v.shape[0] > 50000:      # nssens3D: every 11th site and the SHA-256 of the whole array
                    res[f"{prog}.{k}/s11"] = v[::11].copy()
                    res[f"{prog}.{k}/sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(v + 0.0).tobytes()).digest(), dtype=np.uint8)
                else:
                    res[f"{prog}.{k}"] = v
            print(prog, {k: v.shape for k, v in arrs.items()}, flush=True)
    np.savez_compressed(os.path.join(HERE, "dropin_more.npz"), **res)
    print(os.path.getsize(os.path.join(HERE, "dropin_more.npz")), "bytes")


if __name__ == "__main__":
    sys.exit(main())
