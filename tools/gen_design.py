"""write the design / parameter files tests/dropin/heatsink_dump.cpp reads (same closed-form design as tests/heatsink_case.py)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import heatsink_case as H
from helpers import gcoords

dim, lx, ly, lz, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
os.makedirs(d, exist_ok=True)
size = (lx, ly, lz)
p = H.params(dim, size)
for name, a in zip(("alpha", "kappa", "dads", "dkds"), H.design_fields(p, *gcoords(*size))):
    np.ascontiguousarray(a, dtype=np.float64).tofile(os.path.join(d, name + ".bin"))
np.array([p["nu"], p["gx"], p["gy"], p["gz"], p["tem0"], p["qn0"], p["L"]]).tofile(os.path.join(d, "params.bin"))
