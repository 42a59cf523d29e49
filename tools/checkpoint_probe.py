"""Timing probe for the checkpoint-recompute state store (panslbm2_b200/transient.py) on the transient heatsink loops of BASELINE
configs[4] at the production size: the same sweep with every step stored (every = 1, the reference's scheme) and with every K-th
step stored.  Prints one JSON line per setting: ms per forward step, ms per adjoint step (adjoint collide + sensitivity + the
recomputed forward steps), bytes of states and checkpoints held, and whether dfdss is bit-identical to the store-all run.
    python tools/checkpoint_probe.py [--dims 81,161,81] [--nt 200] [--every 1,8,14,32]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="81,161,81")
    ap.add_argument("--nt", type=int, default=200)
    ap.add_argument("--every", default="1,8,14,32")
    a = ap.parse_args()
    from transient_case import run_transient_cuda
    import panslbm2_b200 as pl
    size = tuple(int(v) for v in a.dims.split(","))
    base = None
    for every in [int(v) for v in a.every.split(",")]:
        for rep in range(2):       # the second run reuses the allocator's warm pools
            st = {}
            res = run_transient_cuda(size, a.nt, every, st)
        if base is None:
            base = res["dfdss"]
        line = {"workload": f"transient heatsink3D loops {size[0]}x{size[1]}x{size[2]}, nt = {a.nt}", "every": every, **{k: st[k] for k in sorted(st)},
                "dfdss_bit_identical_to_first_setting": bool(np.array_equal(base, res["dfdss"])), "max_abs_dfdss": float(np.max(np.abs(res["dfdss"])))}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
