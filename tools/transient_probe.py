"""BASELINE configs[4] (production/heatsink3D_transient.cpp) at its committed size 81 x 161 x 81: the transient forward loop
(one set of macroscopic arrays + one thermal snapshot stored per step) and the time-reversed adjoint loop with a sensitivity
accumulation per step, through the drop-in C++ surface on the GPU, and the same source over the reference headers on the host
cores (oracle/_ref/transient_ref3).  Prints one JSON line.
    python tools/transient_probe.py [nt] [budget_mb] [cpu_nt]"""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import heatsink_case as H
from helpers import gcoords

# PROBE_PE="mx,my,mz": the same on a PE grid, one process per GPU under tools/mpiexec_b200 (strong scaling of the 81 x 161 x 81 domain)
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 200
budget = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cpu_nt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
size = (81, 161, 81)
env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
lib = os.path.join(ROOT, "panslbm2_b200")
out = {"workload": f"transient heatsink loops 81x161x81, nt={nt} stored steps ({23*8*size[0]*size[1]*size[2]*nt/1e9:.1f} GB of states)" +
                   (f", PE grid {os.environ['PROBE_PE']} (one process per GPU)" if os.environ.get("PROBE_PE") else "")}
with tempfile.TemporaryDirectory() as d:
    exe = os.path.join(d, "td3")
    pe = os.environ.get("PROBE_PE", "")
    nranks = eval(pe.replace(",", "*")) if pe else 1
    subprocess.check_call(["g++", "-O2", "-mavx", "-ffp-contract=off", "-w", "-DTRANSIENT_DIM=3", "-DPANSLBM_B200_DROPIN"] + (["-DTRANSIENT_MPI"] if pe else []) + ["-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(lib, "src"), os.path.join(ROOT, "tests", "dropin", "transient_dump.cpp"), "-o", exe,
                           "-L" + lib, "-lpanslbm_b200", "-Wl,-rpath," + lib], env=env)
    p = H.params(3, size)
    for name, a in zip(("alpha", "kappa", "dads", "dkds"), H.design_fields(p, *gcoords(*size))):
        np.ascontiguousarray(a, dtype=np.float64).tofile(os.path.join(d, name + ".bin"))
    np.array([p["nu"], p["gx"], p["gy"], p["gz"], p["tem0"], p["qn0"], p["L"]]).tofile(os.path.join(d, "params.bin"))

    def run(cmd, extra_env=None):
        e = dict(env); e.update(extra_env or {})
        if pe and cmd[0] == exe:
            e.update(TRANSIENT_PE=pe, PANSLBM_RDV_DIR=d)
            cmd = [os.path.join(ROOT, "tools", "mpiexec_b200"), "-n", str(nranks)] + cmd
        r = subprocess.run(cmd, capture_output=True, text=True, env=e)
        sys.stderr.write(r.stderr)
        if r.returncode:
            return {"error": (r.stdout + r.stderr)[-600:]}
        m = re.search(r"forward (\d+) steps ([\d.]+) ms/step ([\d.]+) MLUPS \| adjoint\+sensitivity ([\d.]+) ms/step ([\d.]+) MLUPS", r.stdout)
        res = {"forward_ms_per_step": float(m.group(2)), "forward_mlups": float(m.group(3)), "adjoint_ms_per_step": float(m.group(4)), "adjoint_mlups": float(m.group(5))}
        m = re.search(r"fused steps (\d+), calls one by one (\d+), uploads (\d+), downloads (\d+)", r.stdout)
        if m:
            res.update(fused_steps=int(m.group(1)), single_calls=int(m.group(2)), uploads=int(m.group(3)), downloads=int(m.group(4)))
        m = re.search(r"spilled (\d+) mirrors, restored (\d+), device bytes now (\d+), peak (\d+)", r.stdout)
        if m:
            res.update(spilled=int(m.group(1)), restored=int(m.group(2)), peak_device_gb=int(m.group(4))/1e9)
        return res
    out["gpu"] = run([exe, "3", *map(str, size), str(nt), d])      # a first optimisation iteration: pays for the first-touch allocation of the store
    iters = int(os.environ.get("PROBE_ITERATIONS", "2"))
    if iters > 1:
        out[f"gpu_iteration_{iters}"] = run([exe, "3", *map(str, size), str(nt), d, str(iters)])      # a later one (the drivers run nitr = 500 over the same arrays)
    if budget:
        out[f"gpu_budget_{budget}MB"] = run([exe, "3", *map(str, size), str(nt), d], {"PANSLBM_B200_DEVICE_BUDGET_MB": str(budget)})
    ref = os.path.join(ROOT, "oracle", "_ref", "transient_ref3")
    if cpu_nt and os.path.exists(ref):
        out["cpu_reference"] = run([ref, "3", *map(str, size), str(cpu_nt), d])
        out["cpu_reference"]["cores"] = os.cpu_count()
        out["cpu_reference"]["nt"] = cpu_nt
print(json.dumps(out))
