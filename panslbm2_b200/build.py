"""Build recipe for libpanslbm_b200.so (hand-written CUDA for sm_100a) — in-tree, explicit nvcc.

    python -m panslbm2_b200.build            # (re)build if sources are newer than the library

-fmad=false: the reference is built with -mavx only (no FMA, README.md:23-25); contracting a*b+c on the GPU would
change the last bits of every population.  The sweep is HBM-bound, so the extra instruction issue is hidden.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpanslbm_b200.so")
SOURCES = ["panslbm_api.cu", "panslbm_host.cpp"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libpanslbm_b200.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "panslbm_c.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    tmp = LIB + ".tmp"       # replaced atomically: a snapshot of the tree (gpurun) never sees a half-written library
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)   # this image exports CC=/opt/gcc/bin/gcc; let nvcc use the system g++
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libpanslbm_b200.so")
    os.replace(tmp, LIB)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
