"""Filter parity cases shared by the fixture generator (reference build) and the GPU test (drop-in headers): deterministic
inputs, the drivers' design-box weight or the headers' default cone weight."""
import numpy as np

# tag: (dim, (lx, ly, lz), R, beta, box or None)
CASES = {
    "hs3d_box": (3, (13, 11, 9), 2.4, 2.0, (10, 9, 7)),     # production/heatsink3D.cpp:44, 87-93 (R = 2.4, design box)
    "hs2d_box": (2, (23, 19, 1), 2.4, 4.0, (18, 15, 1)),     # production/heatsink.cpp:83-89
    "cone3d": (3, (9, 10, 8), 1.8, 1.0, None),               # test/heavisidefilter.cpp:30-31 (R = 1.8, beta = 1)
    "cone2d_r3": (2, (17, 12, 1), 3.0, 8.0, None),           # integer radius: pairs at distance == R carry weight 0
}


def inputs(tag):
    dim, size, R, beta, box = CASES[tag]
    n = size[0]*size[1]*size[2]
    rs = np.random.RandomState(abs(hash(tag)) % 2**31 if False else sum(ord(c) for c in tag))
    v = rs.uniform(0.0, 1.0, n)
    d = rs.uniform(-1.0, 1.0, n)
    return v, d
