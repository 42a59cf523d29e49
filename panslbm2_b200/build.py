"""Build recipe for libpanslbm_b200.so (hand-written CUDA for sm_100a) — in-tree, explicit nvcc.

    python -m panslbm2_b200.build            # (re)build what is older than its sources
    python -m panslbm2_b200.build --force -v

-fmad=false: the reference is built with -mavx only (no FMA, README.md:23-25); contracting a*b+c on the GPU would
change the last bits of every population.  The sweep is HBM-bound, so the extra instruction issue is hidden.

The per-model kernels (k_collide / k_fused / k_shell / k_tubes for each lattice and each of the fourteen Macro*Collide* models,
three pass modes each) are compiled as one translation unit per (lattice, model) pair — csrc/lbm_model_inst.cu with
-DPLI_DIM / -DPLI_MODEL — in parallel; objects land in build/obj/ (git-ignored), the library next to this file.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# PANSLBM_BUILD_TAG=<tag> PANSLBM_BUILD_FLAGS="-D..." builds an A/B variant libpanslbm_b200_<tag>.so beside the library
# (loaded instead of it when PANSLBM_LIB_TAG=<tag> is set, _lib.py); tuning experiments only
TAG = os.environ.get("PANSLBM_BUILD_TAG", "")
OBJ = os.path.join(ROOT, "build", "obj" + ("_" + TAG if TAG else ""))
LIB = os.path.join(HERE, "libpanslbm_b200" + ("_" + TAG if TAG else "") + ".so")
PAIRS = [(2, m) for m in range(1, 15)] + [(3, m) for m in range(1, 12)]      # models 12 (mass flow), 13, 14 (NSin) exist for D2Q9 only
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC"] + \
    os.environ.get("PANSLBM_BUILD_FLAGS", "").split()


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libpanslbm_b200.so cannot be built (there is no CPU fallback)")


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "panslbm_c.h"), os.path.abspath(__file__)]


def _units():
    """(object, source, extra flags)"""
    u = [(os.path.join(OBJ, "panslbm_api.o"), os.path.join(CSRC, "panslbm_api.cu"), []),
         (os.path.join(OBJ, "panslbm_host.o"), os.path.join(CSRC, "panslbm_host.cpp"), [])]
    # the two hot pairs first: they take longest
    for d, m in sorted(PAIRS, key=lambda p: p not in ((3, 7), (3, 11))):
        u.append((os.path.join(OBJ, f"model_{d}_{m}.o"), os.path.join(CSRC, "lbm_model_inst.cu"), [f"-DPLI_DIM={d}", f"-DPLI_MODEL={m}"]))
    return u


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps() if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)   # this image exports CC=/opt/gcc/bin/gcc; let nvcc use the system g++
    nvcc = _nvcc()
    newest = max(os.path.getmtime(d) for d in _deps() if os.path.isfile(d))
    log = []

    def compile_one(unit):
        obj, src, extra = unit
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
            return
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + extra + ["-c", src, "-o", obj + ".tmp.o"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        os.replace(obj + ".tmp.o", obj)
        if verbose:
            log.append(r.stderr)

    with ThreadPoolExecutor(max_workers=max(1, min(len(os.sched_getaffinity(0)), 12))) as ex:
        list(ex.map(compile_one, _units()))
    tmp = LIB + ".tmp"       # replaced atomically: a snapshot of the tree (gpurun) never sees a half-written library
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + [u[0] for u in _units()], capture_output=True, text=True, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libpanslbm_b200.so")
    os.replace(tmp, LIB)
    if verbose:
        sys.stderr.write("".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
