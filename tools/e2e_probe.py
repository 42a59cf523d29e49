"""where the end-to-end time of bench.py goes (host phases), for tuning only"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import panslbm2_b200 as pl
from panslbm2_b200 import _lib, api
import bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 352
K = 20
L = _lib.lib()
t = [time.perf_counter()]
def lap(name):
    torch.cuda.synchronize(); L.pl_synchronize()
    t.append(time.perf_counter()); print(f"{name:28s} {1e3*(t[-1]-t[-2]):9.1f} ms", flush=True)
sw = bench.HeatsinkSweep(pl, api, (S, S, S)); lap("construct")
sw.upload_design(); lap("upload design (pageable)")
sw.init_forward(); lap("init_forward + bake plan")
sw.fplan.advance(3, end_streamed=False); lap("3 fwd steps")
sw.fplan.advance(0, end_streamed=True); lap("close step")
sw.init_adjoint(); lap("init_adjoint + bake plan")
sw.aplan.advance(3, end_streamed=False); lap("3 adj steps")
hdesign = [torch.from_numpy(a).pin_memory() for a in sw.host_design]
hout = [torch.empty(sw.n, dtype=torch.float64).pin_memory() for _ in range(2)]; lap("pin")
for rep in range(2):
    sw.upload_design(hdesign); lap("e2e: upload pinned")
    sw.init_forward(); lap("e2e: init_forward")
    sw.fplan.advance(K, end_streamed=True); lap(f"e2e: {K} fwd")
    sw.init_adjoint(); lap("e2e: init_adjoint")
    sw.aplan.advance(K, end_streamed=True); lap(f"e2e: {K} adj")
    sw.sensitivity(); lap("e2e: sensitivity")
    _lib.check(L.pl_array_download(hout[0].data_ptr(), sw.dfdss.ptr, sw.n)); _lib.check(L.pl_array_download(hout[1].data_ptr(), sw.A["tem"].ptr, sw.n)); lap("e2e: download")
