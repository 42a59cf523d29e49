"""A/B probe for the launch-bound regime: bench.py's small-domain sub-lines (cavity 101^2, heatsink 141x161 forward / adjoint, D2Q9) and one
rank's block of configs[3] on 2x2x2 (41x81x41, D3Q15 NS+AD) under the schedule knobs of DESIGN.md §3 (PANSLBM_GRAPH, PANSLBM_XGHOST, ...),
which are read once per process: run it once per setting.  Prints one JSON line.
    PANSLBM_GRAPH=1 python tools/small_domain_probe.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import bench
    import panslbm2_b200 as pl
    from panslbm2_b200 import _lib, api
    _lib.check(_lib.lib().pl_set_device(0))
    out = {"knobs": {k: v for k, v in os.environ.items() if k.startswith("PANSLBM_")}}
    sd = bench.small_domains(pl, api, torch, with_reference=False)
    out["cavity2d_us"] = sd["cavity2d_101x101"]["us_per_step"]
    out["heatsink2d_fwd_us"] = sd["heatsink2d_141x161"]["forward_us_per_step"]
    out["heatsink2d_adj_us"] = sd["heatsink2d_141x161"]["adjoint_us_per_step"]
    for tag, size in (("41x81x41", (41, 81, 41)), ("81x161x81", (81, 161, 81))):
        sw = bench.HeatsinkSweep(pl, api, size)
        sw.upload_design()
        sw.init_forward()
        sw.fplan.advance(50, end_streamed=False, save_last=bench.SAVE_LAST)
        K = 400
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize(); e[0].record()
        sw.fplan.advance(K, end_streamed=False, save_last=bench.SAVE_LAST)
        e[1].record(); torch.cuda.synchronize()
        sw.fplan.advance(0, end_streamed=True)
        sw.init_adjoint()
        sw.aplan.advance(50, end_streamed=False, save_last=bench.SAVE_LAST)
        torch.cuda.synchronize(); e[2].record()
        sw.aplan.advance(K, end_streamed=False, save_last=bench.SAVE_LAST)
        e[3].record(); torch.cuda.synchronize()
        uf, ua = 1e3*e[0].elapsed_time(e[1])/K, 1e3*e[2].elapsed_time(e[3])/K
        out[f"heatsink3d_{tag}_fwd_us"], out[f"heatsink3d_{tag}_adj_us"] = uf, ua
        out[f"heatsink3d_{tag}_mlups"] = 2*sw.n/(uf + ua)
        del sw
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
