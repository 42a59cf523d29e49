"""Fixtures for the GPU filters from the REFERENCE's own DensityFilter / HeavisideFilter (oracle/_ref, built from the unmodified
headers):  make -C oracle ref && python tests/golden/make_filter_golden.py  -> tests/golden/filters.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
import filter_case as FC  # noqa: E402

out = {}
for tag, (dim, size, R, beta, box) in FC.CASES.items():
    ref = O.Backend("ref", dim)
    l = ref.lattice(*size)
    v, d = FC.inputs(tag)
    b = box or (0, 0, 0)
    for mode, name in enumerate(("fv", "rho", "dfds")):
        res = np.zeros(l.nxyz)
        ref._call("filter", l, mode, float(R), float(beta), v, d if mode == 2 else None, res, *b)
        out[f"{tag}/{name}"] = res
    l.free()
np.savez_compressed(os.path.join(HERE, "filters.npz"), **out)
print({k: (v.shape, float(v.min()), float(v.max())) for k, v in out.items()})
