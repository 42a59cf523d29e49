#!/usr/bin/env python
"""bench.py — MLUPS of the fused D3Q15 stream+collide sweep on B200 (BASELINE.json configs[2]:
test/cavityflow3D.cpp scaled to 512^3), with the HBM roofline of the dominant kernel and the reference's own
CPU path timed beside it.

    python bench.py --gpus 1 --steps 50 --warmup 5            # our arm
    python bench.py --impl reference --steps 5 --warmup 1     # reference arm (oracle/_ref on the host cores)

One "step" = one lattice update of the whole domain: Stream + wall/lid closures + SmoothCorner + MacroCollide,
executed as one fused pass (pl_plan_advance).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG_NS_SAVE = 272.0   # bytes / lattice update: 15 pops read + 15 written + rho,ux,uy,uz written
METRIC = "MLUPS"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples if len(s) > 2 + k)]
        return {"sm_mhz": sm[len(sm)//2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(self.samples)}


def cavity_plan(pl, api, pf, rho, u, nu=0.1, u0=0.1, theta=90.0):
    """record the loop body of test/cavityflow3D.cpp:48-58"""
    import math
    import numpy as np
    lx, ly, lz = pf.lx, pf.ly, pf.lz
    wall = lambda i, j, k: np.where((i == 0) | (i == lx - 1) | (j == 0) | (j == ly - 1) | (k == 0), 1, 0)
    lid = lambda i, j, k: k == lz - 1
    uvals = [lambda i, j, k: u0*math.cos(theta*math.pi/180.0), lambda i, j, k: u0*math.sin(theta*math.pi/180.0), lambda i, j, k: 0.0]
    plan = pl.StepPlan(pf)
    plan.set_collide(pl.collide_args(api.M_NS_COLLIDE, True, nu, rho=rho, ux=u[0], uy=u[1], uz=u[2]))
    plan.add_bounce(pf, wall)
    plan.add_closure(pf, api.BC_NS_SET_U, lid, uvals)
    plan.set_smooth_corner(True).finalize()
    return plan


def cpu_reference(size, steps, warmup, threads=None):
    """reference's own OpenMP+AVX path (oracle/_ref) on the host cores; falls back to the C port when _ref is absent"""
    from oracle import oracle as O
    kind = "reference" if O.have_ref(3) else "port"
    be = O.Backend("ref" if kind == "reference" else "orc", 3)
    if kind == "reference":
        cores = be.lib.ref_max_threads()
        if threads:
            be.lib.ref_set_threads(int(threads)); cores = int(threads)
    else:
        cores = os.cpu_count() or 1
    import numpy as np
    n = size**3
    m = [np.zeros(n) for _ in range(4)]
    sec = be.time_cavity3d(size, size, size, steps, warmup, *m)
    return {"value": n*steps/sec/1e6, "unit": "MLUPS", "cores": int(cores), "kind": kind,
            "sample": f"cavityflow3D D3Q15 NS {size}^3, {steps} steps after {warmup} warm-up, {sec:.2f} s"}, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: probe one step at 96^3, then pick the largest cube <= 256 whose (steps+warmup) fit in ~90 s
    probe, psec = cpu_reference(96, 1, 1)
    rate = probe["value"]*1e6   # sites/s
    budget = 90.0
    size = 96
    for s in (128, 160, 192, 224, 256):
        if s**3*(args.steps + args.warmup)/rate <= budget:
            size = s
    cb, sec = cpu_reference(size, args.steps, args.warmup)
    line = {"metric": METRIC, "value": cb["value"], "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3*sec/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "test/cavityflow3D.cpp D3Q15 NS lid-driven cavity (BASELINE configs[2]), CPU sample " + f"{size}^3",
                       "global_sites": size**3, "parallelism": "OpenMP+AVX host threads"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import panslbm2_b200 as pl
    from panslbm2_b200 import _lib, api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().pl_set_device(local))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    S = args.size
    pf = pl.D3Q15(S, S, S)
    N = pf.nxyz
    rho = pl.DeviceArray(N, 1.0)
    u = [pl.DeviceArray(N, 0.0) for _ in range(3)]
    pl.NS.InitialCondition(pf, rho, *u)
    plan = cavity_plan(pl, api, pf, rho, u)
    L = _lib.lib()

    # ---- device-resident throughput ----------------------------------------------------------------
    plan.advance(args.warmup, end_streamed=False)
    barrier()
    L.pl_launch_count_reset()
    sampler = ClockSampler(local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.pl_plan_profile(plan._h, 1)
    e0.record()
    plan.advance(args.steps, end_streamed=False)
    e1.record()
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    launches = int(L.pl_launch_count())
    import ctypes as C
    kms, kn, ksites = C.c_double(0), C.c_int(0), C.c_longlong(0)
    L.pl_plan_profile_read(plan._h, C.byref(kms), C.byref(kn), C.byref(ksites))
    L.pl_plan_profile(plan._h, 0)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world*N*args.steps/(ms*1e-3)/1e6

    # ---- end to end through the public API with host buffers ---------------------------------------------
    # what test/cavityflow3D.cpp does around its loop: fields start on the host, results are read on the host.
    hrho = torch.ones(N, dtype=torch.float64).pin_memory()
    hu = [torch.zeros(N, dtype=torch.float64).pin_memory() for _ in range(3)]
    barrier()
    t0 = time.perf_counter()
    f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
    f0.record()
    for d, h in zip([rho] + u, [hrho] + hu):
        _lib.check(L.pl_array_upload(d.ptr, h.data_ptr(), N))
    pl.NS.InitialCondition(pf, rho, *u)
    plan2 = cavity_plan(pl, api, pf, rho, u)
    plan2.advance(args.steps, end_streamed=True)
    for d, h in zip([rho] + u, [hrho] + hu):
        _lib.check(L.pl_array_download(h.data_ptr(), d.ptr, N))
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world*N*args.steps/(e2e_ms*1e-3)/1e6
    checksum = float(hu[0].abs().max())

    if rank != 0:
        return
    peak, peak_src = peaks()
    k_avg_ms = kms.value/max(kn.value, 1)
    achieved = B_ALG_NS_SAVE*ksites.value/max(kn.value, 1)/(k_avg_ms*1e-3)/1e9 if kn.value else None
    line = {
        "metric": METRIC, "value": value, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"test/cavityflow3D.cpp D3Q15 NS lid-driven cavity scaled to {S}^3 per GPU (BASELINE configs[2])",
                   "global_sites": world*N, "sites_per_gpu": N, "bytes_per_site_update": B_ALG_NS_SAVE,
                   "parallelism": "1 GPU" if world == 1 else f"{world} independent replicas (halo exchange not wired into bench yet)",
                   "l2": "two 16 GB population buffers per GPU >> 126 MB L2, no flush needed"},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "MLUPS", "h2d_bytes_per_step": 4*N*8/args.steps, "d2h_bytes_per_step": 4*N*8/args.steps,
                "note": "host rho,u -> H2D -> InitialCondition -> plan -> steps -> D2H rho,u (bytes amortised per step)", "max_abs_ux": checksum},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved/peak if achieved else None),
                     "traffic": None, "kernel": "k_fused<3,1>", "avg_kernel_ms": k_avg_ms, "peak_source": peak_src,
                     "kernel_share_of_step": (kms.value/ms if ms else None)},
    }
    if world == 1 and not args.no_cpu:
        try:
            cb, _ = cpu_reference(args.cpu_size, args.cpu_steps, 1)
            line["cpu_baseline"] = cb
        except Exception as ex:   # the checker must never take the bench down
            line["cpu_baseline"] = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--cpu-size", type=int, default=160)
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
