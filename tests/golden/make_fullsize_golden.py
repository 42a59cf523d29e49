"""Full-size fixture: the heatsink3D iteration of tests/heatsink_case.py at the production size of production/heatsink3D.cpp:42
(81 x 161 x 81 = 1 056 321 sites: one scalar-tail site) on the REFERENCE build (oracle/_ref), 2000 forward + 2000 adjoint steps (SURVEY §8d cfg 2/4).
Stores sha256 digests and 1-in-997 samples of every field (the fields themselves are ~8 MB each).
    make -C oracle ref && python tests/golden/make_fullsize_golden.py      (about a quarter of an hour on 8 cores)
With the argument 2: the 2-D twin at the size of production/heatsink.cpp:41 (141 x 161 = 22 701 sites, BASELINE configs[1]),
2000 forward + 2000 adjoint steps -> heatsink2d_fullsize.npz."""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
import heatsink_case as H  # noqa: E402

DIM = int(sys.argv[1]) if len(sys.argv) > 1 else 3
SIZE, NT = ((81, 161, 81), 2000) if DIM == 3 else ((141, 161, 1), 2000)
NAME = "heatsink_fullsize.npz" if DIM == 3 else "heatsink2d_fullsize.npz"
t0 = time.time()
r = H.run_oplevel(O.Backend("ref", DIM), DIM, SIZE, NT)
out = {"shape": np.array(list(SIZE) + [NT])}
for k, a in r.items():
    if k in ("gsnap", "igsnap"):
        continue
    a = a + 0.0
    out[f"{k}/s997"] = a[::997].copy()
    out[f"{k}/sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
np.savez_compressed(os.path.join(HERE, NAME), **out)
print("wrote %s in %.0f s" % (NAME, time.time() - t0), sorted(k for k in out if k.endswith("sha")))
