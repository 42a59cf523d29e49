"""TEST INFRASTRUCTURE: builds tests/hostmath/libhostmath.so — the product's site-local math (csrc/*.cuh) compiled as
host C++ with FMA contraction off, see hostmath.cpp."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhostmath.so")
CSRC = os.path.join(HERE, "..", "..", "panslbm2_b200", "csrc")


def build(force=False):
    deps = [os.path.join(HERE, "hostmath.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(HERE, "hostmath.cpp"), "-o", SO], env=env)
    return SO


if __name__ == "__main__":
    print(build(force=True))
