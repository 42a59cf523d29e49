#!/usr/bin/env python
"""bench.py — MLUPS (fp64, forward + adjoint) of the D3Q15 NS+AD lattice-Boltzmann sweep on B200, with the HBM roofline
of the dominant kernel and the reference's own CPU path timed beside it.

    python bench.py --gpus 1 --steps 30 --warmup 3            # our arm
    python bench.py --impl reference --steps 5 --warmup 1     # reference arm (oracle/_ref on the host cores)

Workload: the forward ("Direct analyse") and adjoint ("Inverse analyse") time loops of production/heatsink3D.cpp
(:148-184, :191-224; the physics of BASELINE configs[3]) on a synthetic S^3 block per GPU (the synthetic-domain scaling of
configs[2]), grey closed-form design, convergence `break` disabled.  One "step" = one forward lattice update
(AD::MacroBrinkmanCollideNaturalConvection + 2x Stream + wall/SetT/SetQ closures + 2x SmoothCorner, macros and the g
snapshot saved) plus one adjoint lattice update (AAD::MacroBrinkmanCollideNaturalConvection + 2x iStream + iSetT/iSetQ/
iSetQ(eps)/walls + SmoothCorner), each executed as one fused pass (pl_plan_advance).  MLUPS = 2*sites*steps/seconds.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# algorithmic bytes per lattice update (SURVEY.md §8d / DESIGN.md): every population read once and written once, per-site
# coefficient fields read; the passes whose outputs somebody can look at also write the macros and the thermal snapshot
B_NS, B_NS_SAVE = 240.0, 272.0      # D3Q15 NS: 15r+15w (+ rho,u: 4w)
B_FWD, B_FWD_SAVE = 496.0, 680.0    # D3Q15 NS+AD forward: 30r+30w + alpha,kappa (2r)  (+ rho,u,T,q: 8w + g snapshot: 15w)
B_ADJ, B_ADJ_SAVE = 536.0, 744.0    # D3Q15 NS+AD adjoint: 30r+30w + rho,u,T,alpha,kappa (7r)  (+ ip,iu,im,iT,iq: 11w + ig snapshot: 15w)
# The reference stores the macros and the snapshot on every step (free on a CPU); its drivers look at them every dt = 100 steps
# (Residual, heatsink3D.cpp:152-160) and after the loop.  The timed loops here are ONE such observation interval: the last two
# collides of the K steps store everywhere (both alternating argument sets are then exactly what the reference holds), the
# others only on the closure planes (pl_plan_advance_observed).  With K = 20 that is 10 % storing passes against the drivers' 2 %.
SAVE_LAST = 2
METRIC = "MLUPS"
# mx, my, mz per GPU count: pencils that keep x whole (y/z block faces are contiguous planes and the x walls stay with k_xclose);
# --pe 2,2,2 runs the reference's own 8-rank grid (production/heatsink3D.cpp:35), measured 5 % slower
PE_GRIDS = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (1, 2, 4)}


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
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples if len(s) > 2 + k)]
        return {"sm_mhz": sm[len(sm)//2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------
class HeatsinkSweep:
    """forward + adjoint plans of the heatsink3D loop bodies on one block (tests/heatsink_case.py holds the same sequence
    call by call for the parity tests)"""

    def __init__(self, pl, api, size, peid=0, m=(1, 1, 1), dim=3):
        import heatsink_case as H
        import numpy as np
        self.pl, self.api, self.H, self.dim = pl, api, H, dim
        self.p = p = H.params(dim, size)
        if dim == 3:
            self.f, self.g = pl.D3Q15(*size, peid, *m), pl.D3Q15(*size, peid, *m)
        else:       # production/heatsink.cpp: the same loop bodies on D2Q9
            self.f, self.g = pl.D2Q9(size[0], size[1], peid, m[0], m[1]), pl.D2Q9(size[0], size[1], peid, m[0], m[1])
        self.n = n = self.f.nxyz
        f = self.f

        class _L:
            nx, ny, nz, offx, offy, offz = f.nx, f.ny, f.nz, f.offsetx, f.offsety, f.offsetz
        self.host_design = [np.ascontiguousarray(a) for a in H.design_fields(p, *H.local_coords(_L))]   # alpha, kappa, dads, dkds
        self.host_ss = np.ascontiguousarray(H.design_variable(p, *H.local_coords(_L)), dtype=np.float64)   # the (filtered) design they come from
        self.ss = None
        self.alpha, self.kappa, self.dads, self.dkds = [pl.DeviceArray(n) for _ in range(4)]
        names = H.FWD + H.ADJ + ["uxp", "uyp", "uzp", "qxp", "qyp", "qzp", "iuxp", "iuyp", "iuzp", "iqxp", "iqyp", "iqzp"]
        self.A = {k: pl.DeviceArray(n, 0.0) for k in names}
        self.gsnap, self.igsnap = pl.DeviceArray(n*self.f.nc), pl.DeviceArray(n*self.f.nc)
        self.dfdss = pl.DeviceArray(n, 0.0)
        P3 = H.predicates(p)
        self.P = P3 if dim == 3 else {k: (lambda fn: (lambda i, j: fn(i, j, 0)))(v) for k, v in P3.items()}
        if dim == 2:
            for k in [k for k in self.A if k.rstrip("p").endswith("z")]:
                self.A[k] = None
        self.fplan = self.aplan = None

    def upload_design(self, pinned=None):
        src = pinned if pinned is not None else self.host_design
        for d, h in zip((self.alpha, self.kappa, self.dads, self.dkds), src):
            d.upload(h) if pinned is None else self._up(d, h)

    def design_map(self):
        """alpha, kappa, dads, dkds from the design on the device (heatsink3D.cpp:114-119: pl_design_map, bit-identical to the host loop)"""
        from panslbm2_b200 import _lib
        p = self.p
        _lib.check(_lib.lib().pl_design_map(self.ss.ptr, self.ss.n, float(p["diff_fluid"]), float(p["diff_solid"]), float(p["qg"]),
                                            float(p["alphamax"]/float(p["ly"] - 1)), float(p["qf"]), self.kappa.ptr, self.alpha.ptr, self.dkds.ptr, self.dads.ptr))

    def objective(self):
        """mean temperature of the heat patch (heatsink3D.cpp:227-240) reduced on the device: 8 bytes come back"""
        import math
        Lp = int(math.ceil(self.p["L"]))
        return self.api.box_sum(self.g, self.A["tem"], 0, Lp, 0, 1, 0, Lp if self.dim == 3 else 1)/float(Lp*(Lp if self.dim == 3 else 1))

    def _up(self, d, t):
        from panslbm2_b200 import _lib
        _lib.check(_lib.lib().pl_array_upload(d.ptr, t.data_ptr(), d.n))

    def init_forward(self):
        pl, A = self.pl, self.A
        A["rho"].fill(1.0)
        for k in ("ux", "uy", "uz", "tem"):
            if A[k] is not None:
                A[k].fill(0.0)
        u = [A[k] for k in ("ux", "uy", "uz") if A[k] is not None]
        pl.NS.InitialCondition(self.f, A["rho"], *u)
        pl.AD.InitialCondition(self.g, A["tem"], *u)
        if self.fplan is None:
            self.fplan = self._forward_plan()

    def _forward_plan(self):
        pl, api, p, A, P, f, g = self.pl, self.api, self.p, self.A, self.P, self.f, self.g

        def args(sw):
            s = "p" if sw else ""
            arrs = {k: A[k + s] for k in ("ux", "uy", "uz", "qx", "qy", "qz") if A[k + s] is not None}
            ca = pl.collide_args(api.M_AD_BRINKMAN_NAT_CONV, True, p["nu"], gx=p["gx"], gy=p["gy"], gz=p["gz"], tem0=p["tem0"], rho=A["rho"], tem=A["tem"],
                                 alpha=self.alpha, diffusivity=self.kappa, snapshot=self.gsnap, **arrs)
            return ca, pl.bc_aux(ux=arrs["ux"], uy=arrs["uy"], uz=arrs.get("uz"), diffusivity=self.kappa)
        (c0, a0), (c1, a1) = args(False), args(True)
        plan = pl.StepPlan(f, g).set_collide(c0, c1).set_stream(False)
        plan.add_bounce(f, P["f_wall"])
        plan.add_closure(g, api.BC_AD_SET_T, P["setT"], [P["tem"]], a0, a1)
        plan.add_closure(g, api.BC_AD_SET_Q, P["setQ"], [P["qn"]], a0, a1)
        plan.add_bounce(g, P["g_wall"])
        return plan.set_smooth_corner(True, True).finalize()

    def init_adjoint(self):
        pl, A = self.pl, self.A
        for k in ("ip", "iux", "iuy", "iuz", "item", "iqx", "iqy", "iqz"):
            if A[k] is not None:
                A[k].fill(0.0)
        V = lambda *names: [A[k] for k in names if A[k] is not None]
        pl.ANS.InitialCondition(self.f, *V("ux", "uy", "uz"), A["ip"], *V("iux", "iuy", "iuz"))
        pl.AAD.InitialCondition(self.g, *V("ux", "uy", "uz"), A["item"], *V("iqx", "iqy", "iqz"))
        if self.aplan is None:
            self.aplan = self._adjoint_plan()

    def _adjoint_plan(self):
        pl, api, p, A, P, f, g = self.pl, self.api, self.p, self.A, self.P, self.f, self.g

        def args(sw):
            s = "p" if sw else ""
            arrs = {k: A[k + s] for k in ("iux", "iuy", "iuz", "iqx", "iqy", "iqz") if A[k + s] is not None}
            fixed = {k: A[k] for k in ("rho", "ux", "uy", "uz", "tem", "ip", "imx", "imy", "imz", "item") if A[k] is not None}
            return pl.collide_args(api.M_AAD_NAT_CONV, True, p["nu"], gx=p["gx"], gy=p["gy"], gz=p["gz"], alpha=self.alpha, diffusivity=self.kappa,
                                   snapshot=self.igsnap, **fixed, **arrs)
        aux0 = pl.bc_aux(ux=A["ux"], uy=A["uy"], uz=A["uz"])
        aux1 = pl.bc_aux(ux=A["ux"], uy=A["uy"], uz=A["uz"], eps=1.0)
        plan = pl.StepPlan(f, g).set_collide(args(False), args(True)).set_stream(True)
        plan.add_closure(g, api.BC_AAD_ISET_T, P["setT"], [], aux0, aux0)
        plan.add_closure(g, api.BC_AAD_ISET_Q, P["setQ"], [], aux0, aux0)
        plan.add_closure(g, api.BC_AAD_ISET_Q, P["source"], [], aux1, aux1)
        plan.add_bounce(g, P["g_wall"], inverse=True)
        plan.add_bounce(f, P["f_wall"], inverse=True)
        return plan.set_smooth_corner(True, True).finalize()

    def sensitivity(self):
        pl, A, P = self.pl, self.A, self.P
        self.dfdss.fill(0.0)
        V = lambda *names: [A[k] for k in names if A[k] is not None]
        pl.AAD.SensitivityTemperatureAtHeatSource(self.g, self.dfdss, *V("ux", "uy", "uz"), *V("imx", "imy", "imz"), self.dads, A["tem"], A["item"],
                                                  *V("iqx", "iqy", "iqz"), self.gsnap, self.igsnap, self.kappa, self.dkds, P["qn"], P["source"])


def read_profile(L, plan):
    """[(ms, launches, sites)] of the interior kernel: [0] = passes that store on the closure planes only, [1] = passes that store everywhere"""
    kms, kn, ksites = (C.c_double*2)(), (C.c_int*2)(), (C.c_longlong*2)()
    L.pl_plan_profile_read2(plan._h, kms, kn, ksites)
    return [(kms[c], kn[c], ksites[c]) for c in range(2)]


def ncu_traffic(key, grid_threads):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` captures
    (profiles/r02_ncu_traffic.json, else r01), if it was taken on a grid of this size; else None"""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            k = json.load(open(os.path.join(ROOT, "profiles", name)))["kernels"][key]
            if int(k["grid_threads"]) == int(grid_threads):
                return float(k["dram_bytes_per_launch"])
        except Exception:
            pass
    return None


def roofline(kernel, bytes_per_site, prof, peak, peak_src, step_ms_total, traffic=None):
    kms, kn, ksites = prof
    if not kn:
        return None
    avg_ms = kms/kn
    achieved = bytes_per_site*(ksites/kn)/(avg_ms*1e-3)/1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved/peak, "frac_of_nominal_8000_GBs": achieved/8000.0, "traffic": traffic,
            "algorithmic_bytes_per_launch": bytes_per_site*(ksites/kn), "kernel": kernel,
            "avg_kernel_ms": avg_ms, "sites_per_launch": ksites/kn, "algorithmic_bytes_per_site": bytes_per_site, "peak_source": peak_src,
            "kernel_share_of_timed_region": kms/step_ms_total if step_ms_total else None}


# ---------------------------------------------------------------------------------------------------------
REF_BYTES_PER_SITE = 968      # the reference harness at S^3: 2 x (f0 + f + fnext) + 31 fields + 2 snapshots + alpha, kappa = 121 doubles per site


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference(size, steps, warmup, threads=None):
    """the reference's own OpenMP+AVX forward+adjoint loops (oracle/_ref, built from the unmodified headers) on the host cores.
    Launchers export OMP_NUM_THREADS=1 (torch.distributed.run does): the thread count is set explicitly to every core this
    process may run on."""
    import numpy as np
    import heatsink_case as H
    from oracle import oracle as O
    threads = int(threads or host_threads())
    os.environ["OMP_NUM_THREADS"] = str(threads)
    if not O.have_ref(3):
        # the reference build did not travel: time the C restatement of the same loops instead (kind "port")
        sz = (size, size, size)
        be = O.Backend("orc", 3)
        if hasattr(be.lib, "orc_set_threads"):
            be.lib.orc_set_threads(threads)
        secs = H.time_oplevel(be, sz, int(steps), int(warmup))
        n, tot = size**3, float(sum(secs))
        return {"value": 2*n*steps/tot/1e6, "unit": "MLUPS", "cores": threads, "kind": "port",
                "sample": f"heatsink3D forward+adjoint loops on oracle/lbm_oracle.c (OpenMP), D3Q15 NS+AD {size}^3, {steps}+{steps} steps after {warmup}+{warmup} "
                          f"warm-up, {tot:.2f} s (forward {n*steps/secs[0]/1e6:.1f} / adjoint {n*steps/secs[1]/1e6:.1f} MLUPS)"}, tot
    be = O.Backend("ref", 3)
    be.lib.ref_set_threads(threads)
    cores = int(be.lib.ref_max_threads())
    sz = (size, size, size)
    p = H.params(3, sz)
    alpha, kappa, _, _ = [np.ascontiguousarray(a) for a in H.design_fields(p, *H.gcoords(*sz))]
    secs = np.zeros(2)
    be.time_heatsink(size, size, size, alpha, kappa, p["nu"], p["gx"], p["gy"], p["gz"], p["tem0"], p["qn0"], p["L"], int(steps), int(warmup), secs)
    n = size**3
    tot = float(secs.sum())
    return {"value": 2*n*steps/tot/1e6, "unit": "MLUPS", "cores": cores, "kind": "reference",
            "sample": f"heatsink3D forward+adjoint loops, D3Q15 NS+AD {size}^3, {steps}+{steps} steps after {warmup}+{warmup} warm-up, "
                      f"{tot:.2f} s (forward {n*steps/secs[0]/1e6:.1f} / adjoint {n*steps/secs[1]/1e6:.1f} MLUPS)"}, tot


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    # bounded sample: probe at 64^3, then the GPU arm's own block size if its (steps + warmup) forward+adjoint steps fit in ~100 s
    # and its 968 B/site in 40 % of the free host memory; else the largest smaller cube that does
    probe, _ = cpu_reference(64, 2, 1)
    rate = probe["value"]*1e6
    try:
        import psutil
        free = psutil.virtual_memory().available
    except Exception:
        free = 32 << 30
    size = 64
    for s in sorted({96, 128, 160, 192, 224, 256, 288, 320, args.size}):
        if s <= args.size and 2*s**3*(args.steps + args.warmup)/rate <= 100.0 and REF_BYTES_PER_SITE*s**3 <= 0.4*free:
            size = s
    cb, sec = cpu_reference(size, args.steps, args.warmup)
    same = size == args.size
    line = {"metric": METRIC, "value": cb["value"], "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3*sec/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": f"production/heatsink3D.cpp forward+adjoint time loops (D3Q15 NS+AD), " +
                                   (f"the GPU arm's {size}^3 block" if same else f"CPU sample {size}^3 of the {args.size}^3-per-GPU workload (MLUPS is size-normalised)"),
                       "global_sites": size**3, "parallelism": f"OpenMP+AVX, {cb['cores']} host threads", "same_block_as_gpu_arm": same},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def rel_linf(a, b):
    import numpy as np
    d = float(np.max(np.abs(a - b))) if a.size else 0.0
    return d/max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)


def sweep_fields(sw, nt, save_last=SAVE_LAST):
    """nt forward + nt adjoint fused steps + sensitivity on a HeatsinkSweep; every field a driver can look at afterwards"""
    sw.upload_design()
    sw.init_forward()
    sw.fplan.advance(nt, end_streamed=True, save_last=save_last)
    sw.init_adjoint()
    sw.aplan.advance(nt, end_streamed=True, save_last=save_last)
    sw.sensitivity()
    assert nt % 2 == 0      # an even number of std::swap (heatsink3D.cpp:178-183): the plans' first argument set is the drivers' current one
    out = {k: sw.A[k].to_host() for k in sw.H.FWD + sw.H.ADJ}
    out["dfdss"] = sw.dfdss.to_host()
    out["f.f0"], out["g.f0"] = sw.f.get_populations()[0], sw.g.get_populations()[0]
    return out


def block_of(l, a, size):
    import numpy as np
    lx, ly, lz = size
    return np.asarray(a).reshape(lz, ly, lx)[l.offsetz:l.offsetz + l.nz, l.offsety:l.offsety + l.ny, l.offsetx:l.offsetx + l.nx].reshape(-1)


def parity_decomposed(pl, api, torch, dist, rank, m, world):
    """N > 1: the decomposed path with its NCCL halo exchange against a single-block run of the same global domain on this rank's
    own GPU, before the timed region (blocks of 16 x 12 x 8 sites: multiples of 4, so bit-identical is expected; the tolerance
    north_star allows is 1e-10 on the fields and 1e-8 on the sensitivity)"""
    size, nt = (16*m[0], 12*m[1], 8*m[2]), 8
    dec = HeatsinkSweep(pl, api, size, rank, m)
    got = sweep_fields(dec, nt)
    lat = type("B", (), {k: getattr(dec.f, k) for k in ("offsetx", "offsety", "offsetz", "nx", "ny", "nz")})
    del dec
    one = HeatsinkSweep(pl, api, size, 0, (1, 1, 1))
    want = sweep_fields(one, nt)
    del one
    worst = {k: rel_linf(got[k], block_of(lat, want[k], size)) for k in want}
    t = torch.tensor([max(v for k, v in worst.items() if k != "dfdss"), worst["dfdss"]], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mf, ms = float(t[0]), float(t[1])
    return {"ok": bool(mf <= 1e-10 and ms <= 1e-8), "max_rel": max(mf, ms), "max_rel_fields": mf, "max_rel_dfdss": ms, "bit_identical": mf == 0.0 and ms == 0.0,
            "what": f"heatsink3D {nt}+{nt} fused steps + sensitivity on {size[0]}x{size[1]}x{size[2]} decomposed {m[0]}x{m[1]}x{m[2]} over NCCL vs the same domain as one block on "
                    f"each rank's GPU; {len(worst)} fields incl. populations, max over {world} ranks; tolerance 1e-10 fields / 1e-8 dfdss"}


def parity_golden(pl, api, torch, dist, rank, m, world, size):
    """--config heatsink3d: 81x161x81 on the PE grid, 2000 + 2000 fused steps + sensitivity, against the fixture generated from the
    reference build at that size (tests/golden/heatsink_fullsize.npz: 1-in-997 samples + sha256 of every global field)"""
    import hashlib
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "heatsink_fullsize.npz"))
    lx, ly, lz, nt = [int(v) for v in z["shape"]]
    if (lx, ly, lz) != tuple(size):
        return {"ok": None, "what": "no fixture for this size"}
    sw = HeatsinkSweep(pl, api, size, rank, m)
    got = sweep_fields(sw, nt)
    box = (sw.f.offsetx, sw.f.offsety, sw.f.offsetz, sw.f.nx, sw.f.ny, sw.f.nz)
    del sw
    keys = sorted(got)
    if world > 1:
        gathered = [None]*world if rank == 0 else None
        dist.gather_object((box, {k: got[k] for k in keys}), gathered, dst=0)
    else:
        gathered = [(box, got)]
    if rank != 0:
        return None
    worst, exact = {}, True
    for k in keys:
        full = np.zeros((lz, ly, lx))
        for (ox, oy, oz, nx, ny, nz), d in gathered:
            full[oz:oz + nz, oy:oy + ny, ox:ox + nx] = d[k].reshape(nz, ny, nx)
        full = full.reshape(-1) + 0.0
        worst[k] = rel_linf(full[::997], z[f"{k}/s997"])
        exact = exact and hashlib.sha256(np.ascontiguousarray(full).tobytes()).digest() == bytes(z[f"{k}/sha"])
    mf, ms = max(v for k, v in worst.items() if k != "dfdss"), worst["dfdss"]
    return {"ok": bool(mf <= 1e-10 and ms <= 1e-8), "max_rel": max(mf, ms), "max_rel_fields": mf, "max_rel_dfdss": ms, "bit_identical": bool(exact),
            "what": f"heatsink3D {nt}+{nt} fused steps + sensitivity at {lx}x{ly}x{lz} on PE grid {m[0]}x{m[1]}x{m[2]} vs the fixture from the reference build "
                    f"(1-in-997 samples of {len(keys)} global fields, sha256 of each); tolerance 1e-10 fields / 1e-8 dfdss"}


def run_ours(args):
    import torch
    import panslbm2_b200 as pl
    from panslbm2_b200 import _lib, api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().pl_set_device(local))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pl.comm_init_torch()        # NCCL communicator of libpanslbm_b200.so: rank == PEid

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxms(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    L = _lib.lib()
    S, K, W = args.size, args.steps, args.warmup
    save_last = None if args.save_every_step else SAVE_LAST
    # PE grid: z is split first (its faces are contiguous planes), x not at all up to 8 GPUs; --pe overrides
    m = tuple(int(v) for v in args.pe.split(",")) if args.pe else ((2, 2, 2) if args.config == "heatsink3d" and world == 8 else PE_GRIDS.get(world))
    if m is None or m[0]*m[1]*m[2] != world:
        raise SystemExit(f"bench.py: no PE grid defined for {world} GPUs (1, 2, 4, 8; or --pe mx,my,mz)")
    if args.config == "heatsink3d":
        # BASELINE configs[3] as written: production/heatsink3D.cpp:35,42 — 81 x 161 x 81 in total, 2 x 2 x 2 over 8 GPUs
        gsize, scaling = (81, 161, 81), "strong"
    elif args.global_size:
        gsize, scaling = (args.global_size,)*3, "strong"
    else:
        # weak scaling: one S^3 block per GPU, the reference's block decomposition (d3q15.h:29-35) of a (S*mx, S*my, S*mz) domain
        dims = [int(v) for v in args.dims.split(",")] if args.dims else [S, S, S]
        gsize, scaling = (dims[0]*m[0], dims[1]*m[1], dims[2]*m[2]), "weak"

    parity = None
    if world > 1 and not args.no_parity:
        parity = parity_decomposed(pl, api, torch, dist, rank, m, world)
    golden = None
    if args.config == "heatsink3d" and not args.no_parity:
        golden = parity_golden(pl, api, torch, dist, rank, m, world, gsize)

    sw = HeatsinkSweep(pl, api, gsize, rank, m)
    N = sw.n
    NG = gsize[0]*gsize[1]*gsize[2]
    sw.upload_design()

    # ---- device-resident throughput: K forward + K adjoint fused steps -----------------------------------------
    # two timed segments (the adjoint loop starts from the forward loop's final fields, heatsink3D.cpp:191-192): each is W
    # untimed steps, then EXACTLY K steps between a barrier+synchronize on both sides; ms = forward segment + adjoint segment.
    sampler = ClockSampler(local); sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    sw.init_forward()
    sw.fplan.advance(W, end_streamed=False, save_last=save_last)
    L.pl_plan_profile(sw.fplan._h, 1)
    barrier()
    L.pl_launch_count_reset()
    ev[0].record()
    sw.fplan.advance(K, end_streamed=False, save_last=save_last)
    ev[1].record()
    barrier()
    launches = int(L.pl_launch_count())
    sw.fplan.advance(0, end_streamed=True)     # close the last forward step (Stream + closures), as the loop running to nt does
    sw.init_adjoint()
    sw.aplan.advance(W, end_streamed=False, save_last=save_last)
    L.pl_plan_profile(sw.aplan._h, 1)
    barrier()
    L.pl_launch_count_reset()
    ev[2].record()
    sw.aplan.advance(K, end_streamed=False, save_last=save_last)
    ev[3].record()
    barrier()
    launches += int(L.pl_launch_count())
    sampler.stop_flag = True
    fwd_ms, adj_ms = maxms(ev[0].elapsed_time(ev[1])), maxms(ev[2].elapsed_time(ev[3]))
    fprof, aprof = read_profile(L, sw.fplan), read_profile(L, sw.aplan)
    L.pl_plan_profile(sw.fplan._h, 0); L.pl_plan_profile(sw.aplan._h, 0)
    ms = fwd_ms + adj_ms
    value = 2*NG*K/(ms*1e-3)/1e6

    # ---- end to end through the public API with HOST buffers --------------------------------------------------
    # one optimisation-iteration shape (heatsink3D.cpp:114-246): design fields arrive from the host, the loops run, the
    # sensitivity and the temperature field go back to the host.
    # The design variable arrives from the host (one field), its maps alpha, kappa, dalpha/ds, dkappa/ds are evaluated on the device
    # (pl_design_map = heatsink3D.cpp:114-119, bit-identical to the host loop); the objective is reduced on the device
    # (pl_reduce_box_sum = :227-240) and the sensitivity field goes back to the host.
    import numpy as np
    hss = torch.from_numpy(sw.host_ss).pin_memory()
    hout = torch.empty(N, dtype=torch.float64).pin_memory()
    sw.ss = pl.DeviceArray(N)
    want = [a.to_host() for a in (sw.alpha, sw.kappa, sw.dads, sw.dkds)]      # what the host formulas gave (uploaded for the device-resident runs)
    sw.sensitivity()        # warm-up: bakes the heat-source planes once, as the first optimisation iteration of a run does
    sw.objective()          # ... and the first all-reduce of the communicator sets up its channels (N > 1): not part of an iteration
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    sw.ss.upload_async(hss.data_ptr())
    sw.init_forward()       # InitialCondition does not need the design: it runs beside the copy
    api.copy_fence()
    sw.design_map()
    sw.fplan.advance(K, end_streamed=True, save_last=save_last)
    objective = sw.objective()
    sw.init_adjoint()
    sw.aplan.advance(K, end_streamed=True, save_last=save_last)
    sw.sensitivity()
    sw.dfdss.download_async(hout.data_ptr())
    api.copy_wait()
    f1.record()
    barrier()
    e2e_ms = maxms(f0.elapsed_time(f1))
    e2e_value = 2*NG*K/(e2e_ms*1e-3)/1e6
    checks = {"max_abs_dfdss": float(hout.abs().max()), "objective_mean_patch_temperature": float(objective),
              "device_design_map_equals_host_formulas": bool(all(np.array_equal(w, d.to_host()) for w, d in zip(want, (sw.alpha, sw.kappa, sw.dads, sw.dkds))))}

    # ---- secondary sweep: pure NS roofline case (BASELINE configs[2], test/cavityflow3D.cpp scaled) -------------
    extra = {}
    if args.ns_size > 0 and args.config != "heatsink3d":
        del sw
        import gc
        gc.collect()
        extra["ns_cavity"] = ns_cavity(pl, api, L, torch, args.ns_size, max(10, K//2), W, barrier, maxms, world, rank, m, save_last, strong=bool(args.global_size))
    if world == 1 and not args.no_small and args.config != "heatsink3d":
        try:
            extra["small_domains"] = small_domains(pl, api, torch, with_reference=not args.no_cpu)
        except Exception as ex:
            extra["small_domains"] = {"error": repr(ex)}
    if world == 1 and args.filter_size > 0 and args.config != "heatsink3d":
        try:
            extra["filter"] = filter_subline(pl, torch, args.filter_size, with_reference=not args.no_cpu)
        except Exception as ex:
            extra["filter"] = {"error": repr(ex)}

    if world > 1:
        barrier()
        pl.comm_destroy()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = peaks()
    npk = N//4*4        # the grid of k_fused covers the packed sites of the block
    grid = (npk + 255)//256*256
    kf, ka = "k_fused<3,7> (AD::MacroBrinkmanCollideNaturalConvection + Stream x2, fused)", "k_fused<3,11> (AAD::MacroBrinkmanCollideNaturalConvection + iStream x2, fused)"
    # the dominant kernel = the pass that stores on the closure planes only (K - 2 of the K timed launches); the storing pass beside it
    dom = 0 if fprof[0][1] else 1
    rf = roofline(kf + (", every site stores" if dom else ", stores on the closure planes only"), B_FWD_SAVE if dom else B_FWD, fprof[dom], peak, peak_src, fwd_ms, ncu_traffic("k_fused<3,7>" + ("" if dom else "/elided"), grid))
    ra = roofline(ka + (", every site stores" if dom else ", stores on the closure planes only"), B_ADJ_SAVE if dom else B_ADJ, aprof[dom], peak, peak_src, adj_ms, ncu_traffic("k_fused<3,11>" + ("" if dom else "/elided"), grid))
    rfs = roofline(kf + ", every site stores its macros + snapshot", B_FWD_SAVE, fprof[1], peak, peak_src, fwd_ms, ncu_traffic("k_fused<3,7>", grid)) if not dom else None
    ras = roofline(ka + ", every site stores its macros + snapshot", B_ADJ_SAVE, aprof[1], peak, peak_src, adj_ms, ncu_traffic("k_fused<3,11>", grid)) if not dom else None
    policy = "every collide stores its macros + thermal snapshot at every site (as the reference does)" if save_last is None else \
        (f"one observation interval of the drivers (Residual every dt steps, heatsink3D.cpp:152-160): the last {save_last} of the K collides store macros + snapshot at every "
         "site, the others on the closure planes only (pl_plan_advance_observed); every array a driver can look at after the K steps is bit-identical")
    line = {
        "metric": METRIC, "value": value, "unit": "MLUPS", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms/K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"production/heatsink3D.cpp forward+adjoint time loops as written (BASELINE configs[3]: D3Q15 NS+AD, {gsize[0]}x{gsize[1]}x{gsize[2]} in total)"
                                if args.config == "heatsink3d" else
                                f"production/heatsink3D.cpp forward+adjoint time loops (D3Q15 NS+AD, BASELINE configs[3] physics) on a synthetic global domain {gsize[0]}x{gsize[1]}x{gsize[2]} "
                                f"({'one ' + 'x'.join(str(g//q) for g, q in zip(gsize, m)) + ' block per GPU, configs[2] synthetic-domain scaling' if scaling == 'weak' else 'fixed total size'})")
                               + "; 1 step = 1 forward + 1 adjoint lattice update",
                   "global_sites": NG, "sites_per_gpu": N, "lattice_updates_per_step": 2, "save_policy": policy,
                   "parallelism": "1 GPU" if world == 1 else f"block decomposition {m[0]}x{m[1]}x{m[2]} (PE grid of the reference, d3q15.h:29-35), halo exchange "
                                                             "by ncclSend/ncclRecv per step and lattice, overlapped with the interior kernel",
                   "l2": f"population buffers {2*15*8*N/1e9:.1f} GB per GPU (ONE buffer per lattice, updated in place" + ("" if os.environ.get("PANSLBM_INPLACE", "1") != "0" else "; PANSLBM_INPLACE=0: plus two spares") + ") " +
                         (">> 126 MB L2, no flush needed" if 2*15*8*N > 4*126e6 else "(L2-sized: the step is launch-bound, see DESIGN.md)"),
                   "library": os.path.basename(_lib.LIB_PATH)},
        "clocks": sampler.summary(),
        "sweeps": {"forward_mlups": NG*K/(fwd_ms*1e-3)/1e6, "adjoint_mlups": NG*K/(adj_ms*1e-3)/1e6, **extra},
        "e2e": {"value": e2e_value, "unit": "MLUPS", "h2d_bytes_per_step": N*8/K, "d2h_bytes_per_step": (N*8 + 8)/K,
                "note": "pinned host design variable -> H2D (InitialCondition beside it) -> pl_design_map (alpha, kappa, dads, dkds on the device, "
                        "heatsink3D.cpp:114-119) -> K forward -> objective by pl_reduce_box_sum (8 bytes D2H, :227-240) -> adjoint InitialCondition -> "
                        "K adjoint -> SensitivityTemperatureAtHeatSource -> D2H dfdss; copies on the library's copy stream, bytes amortised per step", **checks},
        "gpu_launches": launches,
        "roofline": rf, "roofline_adjoint": ra,
    }
    if rfs:
        line["roofline_storing_pass"], line["roofline_adjoint_storing_pass"] = rfs, ras
    if parity is not None:
        line["parity"] = parity
    if golden is not None:
        line["parity_golden" if parity is not None else "parity"] = golden
    if world == 1 and not args.no_cpu:
        try:
            cb, _ = cpu_reference(args.cpu_size, args.cpu_steps, 1)
            line["cpu_baseline"] = cb
        except Exception as ex:   # the checker must never take the bench down
            line["cpu_baseline"] = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
    print(json.dumps(line), flush=True)


def small_domains(pl, api, torch, with_reference=True):
    """BASELINE configs[0] / configs[1] at their committed sizes: test/cavityflow.cpp (D2Q9 NS, 101 x 101) and the forward + adjoint loops
    of production/heatsink.cpp (D2Q9 NS+AD, 141 x 161) — domains of a few 10^4 sites that live in L2, where a step is bound by kernel
    launches, not by bandwidth.  us per lattice update through the fused plan, with the reference's loops on the host
    cores beside them where oracle/_ref travelled."""
    import math
    import numpy as np
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    # --- cavity 101^2 (test/cavityflow.cpp:31-66)
    lx = ly = 101
    pf = pl.D2Q9(lx, ly)
    n = pf.nxyz
    rho, u = pl.DeviceArray(n, 1.0), [pl.DeviceArray(n, 0.0) for _ in range(2)]
    pl.NS.InitialCondition(pf, rho, *u)
    plan = pl.StepPlan(pf).set_collide(pl.collide_args(api.M_NS_COLLIDE, True, 0.1, rho=rho, ux=u[0], uy=u[1]))
    plan.add_bounce(pf, lambda i, j: np.where((i == 0) | (i == lx - 1) | (j == 0), 1, 0))
    plan.add_closure(pf, api.BC_NS_SET_U, lambda i, j: j == ly - 1, [lambda i, j: 0.1, lambda i, j: 0.0]).set_smooth_corner(True).finalize()
    plan.advance(200, end_streamed=False, save_last=SAVE_LAST)
    K = 2000
    e0, e1 = ev(), ev()
    torch.cuda.synchronize(); e0.record()
    plan.advance(K, end_streamed=False, save_last=SAVE_LAST)
    e1.record(); torch.cuda.synchronize()
    us = 1e3*e0.elapsed_time(e1)/K
    out["cavity2d_101x101"] = {"workload": "test/cavityflow.cpp loop body (D2Q9 NS), 101 x 101", "us_per_step": us, "mlups": n/us}
    # --- heatsink 141 x 161 (production/heatsink.cpp:41, loop bodies :140-214)
    size = (141, 161, 1)
    sw = HeatsinkSweep(pl, api, size, dim=2)
    sw.upload_design()
    sw.init_forward()
    sw.fplan.advance(200, end_streamed=False, save_last=SAVE_LAST)
    K = 1000
    e = [ev() for _ in range(4)]
    torch.cuda.synchronize(); e[0].record()
    sw.fplan.advance(K, end_streamed=False, save_last=SAVE_LAST)
    e[1].record(); torch.cuda.synchronize()
    sw.fplan.advance(0, end_streamed=True)
    sw.init_adjoint()
    sw.aplan.advance(200, end_streamed=False, save_last=SAVE_LAST)
    torch.cuda.synchronize(); e[2].record()
    sw.aplan.advance(K, end_streamed=False, save_last=SAVE_LAST)
    e[3].record(); torch.cuda.synchronize()
    uf, ua = 1e3*e[0].elapsed_time(e[1])/K, 1e3*e[2].elapsed_time(e[3])/K
    out["heatsink2d_141x161"] = {"workload": "production/heatsink.cpp forward + adjoint loop bodies (D2Q9 NS+AD), 141 x 161", "forward_us_per_step": uf,
                                 "adjoint_us_per_step": ua, "mlups": 2*sw.n/(uf + ua)}
    if with_reference:
        try:
            from oracle import oracle as O
            import heatsink_case as H
            if O.have_ref(2):
                be = O.Backend("ref", 2)
                be.lib.ref_set_threads(host_threads())
                p = H.params(2, size)
                alpha, kappa, _, _ = [np.ascontiguousarray(a) for a in H.design_fields(p, *H.gcoords(*size))]
                secs = np.zeros(2)
                be.time_heatsink(size[0], size[1], 1, alpha, kappa, p["nu"], p["gx"], p["gy"], p["gz"], p["tem0"], p["qn0"], p["L"], 2000, 100, secs)
                out["heatsink2d_141x161"]["reference_us_per_step"] = [1e6*float(secs[0])/2000, 1e6*float(secs[1])/2000]
                out["heatsink2d_141x161"]["reference_mlups"] = 2*sw.n*2000/float(secs.sum())/1e6
                out["heatsink2d_141x161"]["reference_threads"] = int(be.lib.ref_max_threads())
        except Exception as ex:
            out["heatsink2d_141x161"]["reference_error"] = repr(ex)
    return out


def filter_subline(pl, torch, S, with_reference=True):
    """HeavisideFilter::GetFilteredVariable with the drivers' design-box weight (production/heatsink3D.cpp:44, 87-103: R = 2.4) on an S^3
    lattice: one GPU kernel per call over the baked weight patterns, next to the reference's serial host loops (SURVEY.md §6:
    1.3 s per call at 128^3, three calls per optimisation iteration)"""
    import numpy as np
    R, beta = 2.4, 2.0
    box = [3*(S - 1)//4 + 1]*3

    def weight(i1, j1, k1, i2, j2, k2):
        inside = (i1 < box[0]) & (j1 < box[1]) & (k1 < box[2]) & (i2 < box[0]) & (j2 < box[1]) & (k2 < box[2])
        cone = (R - np.sqrt((i1 - i2)**2.0 + (j1 - j2)**2.0 + (k1 - k2)**2.0))/R
        return np.where(inside, cone, np.where((i1 == i2) & (j1 == j2) & (k1 == k2), 1.0, 0.0))
    p = pl.D3Q15(S, S, S)
    t0 = time.perf_counter()
    f = pl.ConeFilter(p, R, weight)
    bake_s = time.perf_counter() - t0
    n = S**3
    idx = np.arange(n)
    v = 0.5 + 0.4*np.sin(0.37*(idx % S))*np.cos(0.23*((idx//S) % S))*np.sin(0.31*(idx//(S*S)) + 0.5)
    dv = pl.DeviceArray.from_host(v)
    out = pl.DeviceArray(n)
    for _ in range(3):
        f.heaviside(dv, beta, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        f.heaviside(dv, beta, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    line = {"workload": f"HeavisideFilter::GetFilteredVariable, design-box cone weight R = {R}, {S}^3 (production/heatsink3D.cpp:87-103)", "gpu_ms_per_call": ms,
            "weight_patterns": f.npatterns, "device_table_bytes": f.npatterns*125*8 + 4*n, "bake_seconds_once": bake_s,
            "algorithmic_bytes_per_site": 20.0, "achieved_GBs": 20.0*n/(ms*1e-3)/1e9}
    if with_reference:
        try:
            from oracle import oracle as O
            if O.have_ref(3):
                ref = O.Backend("ref", 3)
                l = ref.lattice(S, S, S)
                res = np.zeros(n)
                t0 = time.perf_counter()
                ref._call("filter", l, 1, float(R), float(beta), np.ascontiguousarray(v), None, res, *box)
                line["reference_ms_per_call"] = 1e3*(time.perf_counter() - t0)
                line["max_abs_diff_vs_reference"] = float(np.max(np.abs(out.to_host() - res)))
                l.free()
        except Exception as ex:
            line["reference_ms_per_call"] = None
            line["reference_error"] = repr(ex)
    return line


def ns_cavity(pl, api, L, torch, S, K, W, barrier, maxms, world, rank=0, m=(1, 1, 1), save_last=SAVE_LAST, strong=False):
    """test/cavityflow3D.cpp:44-59 scaled to S^3 per GPU (S^3 in total with --global-size): NS::MacroCollide(save) + Stream + 5 BARRIER walls + lid SetU + SmoothCorner"""
    import math
    import numpy as np
    GX, GY, GZ = (S, S, S) if strong else (S*m[0], S*m[1], S*m[2])
    pf = pl.D3Q15(GX, GY, GZ, rank, *m)
    N = pf.nxyz
    rho = pl.DeviceArray(N, 1.0)
    u = [pl.DeviceArray(N, 0.0) for _ in range(3)]
    pl.NS.InitialCondition(pf, rho, *u)
    nu, u0, theta = 0.1, 0.1, 90.0
    wall = lambda i, j, k: np.where((i == 0) | (i == GX - 1) | (j == 0) | (j == GY - 1) | (k == 0), 1, 0)
    lid = lambda i, j, k: k == GZ - 1
    uv = [lambda i, j, k: u0*math.cos(theta*math.pi/180.0), lambda i, j, k: u0*math.sin(theta*math.pi/180.0), lambda i, j, k: 0.0]
    plan = pl.StepPlan(pf).set_collide(pl.collide_args(api.M_NS_COLLIDE, True, nu, rho=rho, ux=u[0], uy=u[1], uz=u[2]))
    plan.add_bounce(pf, wall).add_closure(pf, api.BC_NS_SET_U, lid, uv).set_smooth_corner(True).finalize()
    plan.advance(W, end_streamed=False, save_last=save_last)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.pl_plan_profile(plan._h, 1)
    e0.record()
    plan.advance(K, end_streamed=False, save_last=save_last)
    e1.record()
    barrier()
    ms = maxms(e0.elapsed_time(e1))
    prof = read_profile(L, plan)
    peak, src = peaks()
    dom = 0 if prof[0][1] else 1
    grid = (N//4*4 + 255)//256*256
    r = roofline("k_fused<3,1> (NS::MacroCollide + Stream, fused)" + (", every site stores" if dom else ", stores on the closure planes only"), B_NS_SAVE if dom else B_NS, prof[dom], peak, src, ms,
                 ncu_traffic("k_fused<3,1>" + ("" if dom else "/elided"), grid))
    out = {"workload": f"test/cavityflow3D.cpp scaled to {GX}x{GY}x{GZ} in total (BASELINE configs[2])", "mlups": GX*GY*GZ*K/(ms*1e-3)/1e6, "ms_per_step": ms/K, "steps": K, "roofline": r}
    if not dom:
        out["roofline_storing_pass"] = roofline("k_fused<3,1>, every site stores rho, u", B_NS_SAVE, prof[1], peak, src, ms, ncu_traffic("k_fused<3,1>", grid))
    return out


def main():
    # stdout carries exactly ONE JSON line: anything a library prints there meanwhile (NCCL's version banner, ...) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=352, help="edge of the cubic block per GPU for the NS+AD forward+adjoint sweep")
    ap.add_argument("--dims", default="", help="lx,ly,lz of the block per GPU instead of --size^3 (e.g. 81,161,81 = production/heatsink3D.cpp:42)")
    ap.add_argument("--pe", default="", help="PE grid mx,my,mz instead of the default for the GPU count (e.g. 2,2,2 = production/heatsink3D.cpp:35)")
    ap.add_argument("--config", default="synthetic", choices=["synthetic", "heatsink3d"],
                    help="heatsink3d = BASELINE configs[3] as written: 81x161x81 in total on the PE grid (2x2x2 on 8 GPUs), strong scaling, checked against the reference fixture")
    ap.add_argument("--global-size", type=int, default=0, help="strong scaling: edge of the cubic GLOBAL domain split over the GPUs (BASELINE.md: 512)")
    ap.add_argument("--ns-size", type=int, default=512, help="edge of the secondary NS cavity sweep (0 = skip)")
    ap.add_argument("--filter-size", type=int, default=128, help="edge of the Heaviside-filter sub-line (0 = skip)")
    ap.add_argument("--save-every-step", action="store_true", help="every collide stores macros + snapshot at every site (the reference's own cadence) instead of the observed policy")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-small", action="store_true", help="skip the configs[0]/[1] sub-lines (2-D domains at their committed sizes)")
    ap.add_argument("--cpu-size", type=int, default=128)
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
