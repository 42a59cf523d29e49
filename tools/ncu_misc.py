"""Workload for the ncu captures of the small kernels that had no profile yet: k_halo_pack (the one-launch pack of the 26 outgoing
messages of a decomposed block), k_init (InitialCondition) and k_design_map (heatsink3D.cpp:114-119 on the device).  Two ranks of a
1 x 1 x 2 PE grid live in this one process (pl_comm_init_loopback) so that a single GPU exercises the halo path.
    ncu --set full -k regex:'k_halo_pack|k_init|k_design_map' ... python tools/ncu_misc.py"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import panslbm2_b200 as pl          # noqa: E402
from panslbm2_b200 import api       # noqa: E402

S = int(os.environ.get("NCU_MISC_SIZE", "192"))
size, m = (S, S, S), (1, 1, 2)
pl.comm_init_loopback(2)
lat = [pl.D3Q15(*size, r, *m) for r in range(2)]
lx, ly, lz = size
wall = lambda i, j, k: np.where((i == 0) | (i == lx - 1) | (j == 0) | (j == ly - 1) | (k == 0), 1, 0)
lid = lambda i, j, k: k == lz - 1
uvals = [lambda i, j, k: 0.0*i, lambda i, j, k: 0.1 + 0.0*i, lambda i, j, k: 0.0*i]
rho = [pl.DeviceArray(l.nxyz, 1.0) for l in lat]
u = [[pl.DeviceArray(l.nxyz, 0.0) for _ in range(3)] for l in lat]
plans = []
for l, r, v in zip(lat, rho, u):
    pl.NS.InitialCondition(l, r, *v)                                    # k_init
    plan = pl.StepPlan(l).set_collide(pl.collide_args(api.M_NS_COLLIDE, True, 0.1, rho=r, ux=v[0], uy=v[1], uz=v[2]))
    plan.add_bounce(l, wall).add_closure(l, api.BC_NS_SET_U, lid, uvals)
    plans.append(plan.set_smooth_corner(True).finalize())
for _ in range(4):
    for plan in plans:
        plan.advance(1, end_streamed=False)                             # k_halo_pack after every pass
ss = pl.DeviceArray.from_host(0.5 + 0.4*np.sin(np.arange(S**3)*1e-3))
out = pl.design_map(ss, 0.1/6.0, 1.0/6.0, 1.0, 1e4/(S - 1), 1e-2)     # k_design_map
pl.synchronize()
print("ok", float(out[1].to_host().max()), math.isfinite(float(rho[0].to_host().sum())))
del plans, lat
pl.comm_destroy()
