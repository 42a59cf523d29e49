"""Generates the committed fixtures from the REFERENCE ITSELF (oracle/_ref, built from the unmodified headers in
/root/reference by oracle/Makefile).  Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures pin both the C restatement (tests/test_oracle_golden.py, CPU) and the CUDA path (tests -m gpu)
on machines where /root/reference and oracle/_ref do not exist.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def digest(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def cavity3d():
    """test/cavityflow3D.cpp call sequence (reference harness ref_time_cavity3d)"""
    ref = O.Backend("ref", 3)
    out = {}
    for tag, (lx, ly, lz, nt) in {"a": (15, 13, 11, 200), "b": (31, 31, 31, 1000)}.items():
        n = lx*ly*lz
        m = [np.zeros(n) for _ in range(4)]
        ref.time_cavity3d(lx, ly, lz, nt, 0, *m)
        out[f"{tag}_shape"] = np.array([lx, ly, lz, nt])
        out[f"{tag}_sha256"] = np.frombuffer(bytes.fromhex(digest(*m)), dtype=np.uint8)
        if n <= 4000:
            for name, a in zip(("rho", "ux", "uy", "uz"), m):
                out[f"{tag}_{name}"] = a
        else:
            for name, a in zip(("rho", "ux", "uy", "uz"), m):
                out[f"{tag}_{name}_s37"] = a[::37].copy()
    np.savez_compressed(os.path.join(HERE, "cavity3d.npz"), **out)


if __name__ == "__main__":
    cavity3d()
    print("wrote", os.listdir(HERE))
