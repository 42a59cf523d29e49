"""Fixtures for tests/test_gpu_nsopt.py: tests/dropin/nsopt_dump.cpp (the loops of production/nsopt.cpp:82-150) built against the
UNMODIFIED reference headers (-I/root/reference/src, README flags + -O2 -ffp-contract=off) and run on the CPU in this container,
twice: with the AVX overloads (-DNSOPT_AVX) and as nsopt.cpp is committed, scalar templates at every site ("<tag>.scalar/...").  The
drop-in headers honour the same macro (pl_set_scalar_order), so each build of the test program is compared with its own fixture,
bit for bit.
    python tests/golden/make_nsopt_golden.py        (needs /root/reference; writes tests/golden/nsopt.npz)
Per case, build and output array: SHA-256 of the raw fp64 bytes and every 5th value."""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
# tag -> (lx, ly, nt, dt); 201 x 201 is the driver's own size (nsopt.cpp:36), 40 401 sites = 1 in the scalar tail
NSOPT_CASES = {"pipe": (201, 201, 600, 100), "pipe_small": (43, 37, 150, 50)}


def main():
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for build, flags in (("avx", ["-DNSOPT_AVX"]), ("scalar", [])):
            exe = os.path.join(d, "nsopt_ref_" + build)
            subprocess.check_call(["g++", "-O2", "-mavx", "-fopenmp", "-ffp-contract=off", "-w", *flags, "-I" + os.path.join(REF, "src"),
                                   os.path.join(os.path.dirname(HERE), "dropin", "nsopt_dump.cpp"), "-o", exe], env=env)
            for tag, (lx, ly, nt, dt) in NSOPT_CASES.items():
                w = os.path.join(d, build + tag)
                os.makedirs(w)
                r = subprocess.run([exe, str(lx), str(ly), str(nt), str(dt), w], capture_output=True, text=True, check=True)
                print(build, tag, r.stdout.strip())
                for f in sorted(os.listdir(w)):
                    if f.endswith(".out"):
                        a = np.fromfile(os.path.join(w, f)) + 0.0
                        key = f"{tag}/{f[:-4]}" if build == "avx" else f"{tag}.scalar/{f[:-4]}"
                        res[key + "/sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
                        res[key + "/s5"] = a if f == "extra.out" else a[::5]
    np.savez_compressed(os.path.join(HERE, "nsopt.npz"), **res)
    print(len(res), "entries", os.path.getsize(os.path.join(HERE, "nsopt.npz")), "bytes")


if __name__ == "__main__":
    sys.exit(main())
