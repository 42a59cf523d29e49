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
    # c = SURVEY §8d cfg 3 parity case: 128^3, 200 steps
    for tag, (lx, ly, lz, nt) in {"a": (15, 13, 11, 200), "b": (31, 31, 31, 1000), "c": (128, 128, 128, 200)}.items():
        n = lx*ly*lz
        m = [np.zeros(n) for _ in range(4)]
        ref.time_cavity3d(lx, ly, lz, nt, 0, *m)
        out[f"{tag}_shape"] = np.array([lx, ly, lz, nt])
        out[f"{tag}_sha256"] = np.frombuffer(bytes.fromhex(digest(*m)), dtype=np.uint8)
        if n <= 4000:
            for name, a in zip(("rho", "ux", "uy", "uz"), m):
                out[f"{tag}_{name}"] = a
        elif n <= 100000:
            for name, a in zip(("rho", "ux", "uy", "uz"), m):
                out[f"{tag}_{name}_s37"] = a[::37].copy()
        else:
            for name, a in zip(("rho", "ux", "uy", "uz"), m):
                out[f"{tag}_{name}_s997"] = a[::997].copy()
                out[f"{tag}_{name}_sha256"] = np.frombuffer(bytes.fromhex(digest(a + 0.0)), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "cavity3d.npz"), **out)


def ops():
    """every op-level scenario of tests/scenarios.py on the reference build -> sha256 digests (values only: -0.0 == +0.0)"""
    import json
    sys.path.insert(0, os.path.dirname(HERE))
    import scenarios as S
    out = {}
    for dim, size in ((2, (7, 5, 1)), (3, (5, 3, 3)), (3, (6, 4, 4))):
        ref = O.Backend("ref", dim)
        tag = f"d{dim}_{size[0]}x{size[1]}x{size[2]}"
        for model in S.FORWARD_MODELS + S.ADJOINT_MODELS:
            if model.endswith("massflow") and dim == 3:
                continue
            out[f"{tag}/collide/{model}"] = digest(*[a + 0.0 for _, a in S.collide(ref, dim, model, size, 3)])
        for kind in S.CLOSURES:
            if kind == "aad_iset_rho" and dim == 3:
                continue
            out[f"{tag}/closure/{kind}"] = digest(*[a + 0.0 for _, a in S.closure(ref, dim, kind, size, 5)])
        for kind in S.SENSITIVITIES:
            out[f"{tag}/sensitivity/{kind}"] = digest(*[a + 0.0 for _, a in S.sensitivity(ref, dim, kind, size, 9)])
        out[f"{tag}/inits"] = digest(*[a + 0.0 for _, a in S.inits(ref, dim, size, 2)])
    json.dump(out, open(os.path.join(HERE, "ops_digests.json"), "w"), indent=1, sort_keys=True)


HEATSINK_CASES = {"hs3d": (3, (13, 17, 9), 40), "hs2d": (2, (23, 29, 1), 40), "hs3d_tail": (3, (7, 9, 5), 15)}


def heatsink():
    """one heatsink optimisation iteration (tests/heatsink_case.py) on the reference build: digests + strided samples"""
    sys.path.insert(0, os.path.dirname(HERE))
    import heatsink_case as H
    out = {}
    for tag, (dim, size, nt) in HEATSINK_CASES.items():
        r = H.run_oplevel(O.Backend("ref", dim), dim, size, nt)
        for k, a in r.items():
            out[f"{tag}/{k}/s5"] = (a + 0.0)[::5].copy()
            out[f"{tag}/{k}/sha"] = np.frombuffer(bytes.fromhex(digest(a + 0.0)), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "heatsink.npz"), **out)


def heatsink_scalar():
    """the same iteration on the reference headers compiled WITHOUT _USE_AVX_DEFINES (oracle/_ref/*_scalar.so: scalar templates at every
    site, the build of production/nsopt.cpp:2) -> heatsink_scalar.npz: what a drop-in program built without the macro must reproduce
    (pl_set_scalar_order).  dfdss is left out in 3-D: the reference's scalar 3-D SensitivityTemperatureAtHeatSource reads out of bounds
    (_uz / _ig swapped between adjointadvection.h:1536 and :805)."""
    sys.path.insert(0, os.path.dirname(HERE))
    import heatsink_case as H
    out = {}
    for tag in ("hs2d", "hs3d_tail"):
        dim, size, nt = HEATSINK_CASES[tag]
        r = H.run_oplevel(O.Backend("ref_scalar", dim), dim, size, nt)
        for k, a in r.items():
            if dim == 3 and k == "dfdss":
                continue
            out[f"{tag}/{k}/s5"] = (a + 0.0)[::5].copy()
            out[f"{tag}/{k}/sha"] = np.frombuffer(bytes.fromhex(digest(a + 0.0)), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "heatsink_scalar.npz"), **out)


if __name__ == "__main__":
    if "scalar" in sys.argv[1:]:
        heatsink_scalar()
        sys.exit(0)
    cavity3d()
    ops()
    heatsink()
    print("wrote", os.listdir(HERE))
