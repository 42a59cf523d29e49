"""Fixtures for tests/test_gpu_dropin.py: outputs of UNMODIFIED reference programs built with the reference's own headers and
flags (README.md:23-25: g++ -mavx -fopenmp; plus -O2 -ffp-contract=off) and run on the CPU in this container.
    python tests/golden/make_dropin_golden.py          (needs /root/reference; writes tests/golden/dropin.npz)
test/d2q9.cpp, test/d3q15.cpp: stdout (the reference's only known-answer tests: LoadF/StoreF layout round trip).
test/cavityflow3D.cpp: the point data of result/cavity3D_0.vts (rho, u; 6 significant digits as the reference writes them).
test/nsadncsens.cpp: heatsink physics on 71 x 81 with a fixed design, 100 000 forward + 100 000 adjoint steps, sensitivity, Normalize:
every point-data array of result/nsadncsens_0.vts (the best deterministic end-to-end fixture of the reference).
test/heavisidefilter.cpp (hard-wired _USE_MPI_DEFINES): built with the reference headers and, standing in for an MPI library, the
world-of-one path of panslbm2_b200/src/mpi/mpi.h (no device involved); run as `heavisidefilter 1 1 1`; point data v, fv."""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


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
        for prog in ("d2q9", "d3q15", "cavityflow3D", "nsadncsens"):
            exe = os.path.join(d, prog)
            subprocess.check_call(["g++", "-O2", "-mavx", "-fopenmp", "-ffp-contract=off", "-w", os.path.join(REF, "test", prog + ".cpp"), "-o", exe], env=env)
            out = subprocess.run([exe], cwd=d, capture_output=True, text=True, check=True).stdout
            if prog in ("d2q9", "d3q15"):
                res[prog + ".stdout"] = np.frombuffer(out.encode(), dtype=np.uint8)
        for k, v in vts_arrays(os.path.join(d, "result", "cavity3D_0.vts")).items():
            res["cavity3D." + k] = v
        for k, v in vts_arrays(os.path.join(d, "result", "nsadncsens_0.vts")).items():
            res["nsadncsens." + k] = v
        root = os.path.dirname(os.path.dirname(HERE))
        lib = os.path.join(root, "panslbm2_b200")
        exe = os.path.join(d, "heavisidefilter")
        subprocess.check_call(["g++", "-O2", "-mavx", "-fopenmp", "-ffp-contract=off", "-w", "-I" + os.path.join(root, "include"), "-I" + os.path.join(lib, "src", "mpi"),
                               os.path.join(REF, "test", "heavisidefilter.cpp"), "-o", exe, "-L" + lib, "-lpanslbm_b200", "-Wl,-rpath," + lib], env=env)
        subprocess.run([exe, "1", "1", "1"], cwd=d, capture_output=True, text=True, check=True)
        for k, v in vts_arrays(os.path.join(d, "result", "heavisidefilter_0.vts")).items():
            res["heavisidefilter." + k] = v
    np.savez_compressed(os.path.join(HERE, "dropin.npz"), **res)
    print({k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    sys.exit(main())
