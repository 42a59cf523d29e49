"""One rank of the multi-GPU parity check (launched by tests/test_gpu_nccl.py through torch.distributed.run, one process per
GPU): the heatsink3D forward + adjoint loops and the cavity loop on a block-decomposed domain with the halo exchange over
NCCL, compared on every rank with the matching slice of a single-block run of the same global domain on that rank's GPU.
Bit-exact (block sizes are multiples of 4 sites, see test_gpu_decomposed.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def block(l, a, size):
    lx, ly, lz = size
    return np.asarray(a).reshape(lz, ly, lx)[l.offsetz:l.offsetz + l.nz, l.offsety:l.offsety + l.ny, l.offsetx:l.offsetx + l.nx].reshape(-1)


def run(pl, api, size, rank, m, nt):
    import bench
    s = bench.HeatsinkSweep(pl, api, size, rank, m)
    s.upload_design()
    s.init_forward()
    s.fplan.advance(nt//2, end_streamed=False)
    res_mid = pl.Residual(s.A["ux"], s.A["uy"], s.A["uz"], s.A["uxp"], s.A["uyp"], s.A["uzp"], s.n)    # global over the communicator
    s.fplan.advance(nt - nt//2, end_streamed=True)
    s.init_adjoint()
    s.aplan.advance(nt, end_streamed=True)
    s.sensitivity()
    out = {k: v.to_host() for k, v in s.A.items()}
    out["dfdss"] = s.dfdss.to_host()
    f0, f = s.f.get_populations()
    g0, g = s.g.get_populations()
    out["f0"], out["g0"] = f0, g0
    for c in range(1, 15):
        out[f"f{c}"] = f.reshape(-1, 14)[:, c - 1].copy()
        out[f"g{c}"] = g.reshape(-1, 14)[:, c - 1].copy()
    # a call-by-call Stream on the decomposed lattice as well (lazy pack + exchange)
    s.f.Stream(); s.g.iStream()
    out["f0s"] = s.f.get_populations()[1].reshape(-1, 14)[:, 6].copy()
    out["g0s"] = s.g.get_populations()[1].reshape(-1, 14)[:, 10].copy()
    return s.f, out, res_mid


def collectives(rank, world):
    """the communicator entry points the mpi.h shim maps MPI onto (src/mpi/mpi.h): pl_comm_allreduce (<= 4 doubles), pl_comm_allreduce_v
    (any count, f64 / i32, sum / max / min), pl_comm_p2p (grouped sends and receives matched in issue order), pl_reduce_box_sum and
    pl_comm_gather_field — against what MPI would return"""
    import ctypes as C
    import panslbm2_b200 as pl
    from panslbm2_b200 import _lib
    L = _lib.lib()
    ok = True
    v = np.array([rank + 1.0, 10.0*(rank + 1)])
    _lib.check(L.pl_comm_allreduce(v.ctypes.data, 2, 0))
    ok &= np.array_equal(v, [world*(world + 1)/2.0, 10.0*world*(world + 1)/2.0])
    for dtype, arr in ((0, np.arange(1000, dtype=np.float64)*(rank + 1)), (1, (np.arange(333, dtype=np.int32) % 7)*(rank + 1))):
        for op, fn in ((0, lambda a: a*(world*(world + 1)//2)), (1, lambda a: a*world), (2, lambda a: a*1)):
            base = arr/(rank + 1) if dtype == 0 else arr//(rank + 1)
            x = np.ascontiguousarray(arr.copy())
            _lib.check(L.pl_comm_allreduce_v(x.ctypes.data, x.size, dtype, op))
            ok &= np.array_equal(x, fn(base).astype(x.dtype))
    # ring: every rank sends 4096 doubles to rank + 1 and receives from rank - 1 (one group, as MPI_Isend / Irecv / Waitall)
    class P2P(C.Structure):
        _fields_ = [("host", C.c_void_p), ("bytes", C.c_size_t), ("peer", C.c_int), ("is_send", C.c_int)]
    out = np.full(4096, float(rank))
    inp = np.zeros(4096)
    ops = (P2P*2)(P2P(out.ctypes.data, out.nbytes, (rank + 1) % world, 1), P2P(inp.ctypes.data, inp.nbytes, (rank - 1) % world, 0))
    L.pl_comm_p2p.restype, L.pl_comm_p2p.argtypes = C.c_int, [C.c_void_p, C.c_int]
    _lib.check(L.pl_comm_p2p(ops, 2))
    ok &= bool(np.all(inp == float((rank - 1) % world)))
    return bool(ok)


def main():
    import torch
    import torch.distributed as dist
    import panslbm2_b200 as pl
    from panslbm2_b200 import _lib, api
    m = tuple(int(x) for x in sys.argv[1].split("x"))
    size = tuple(int(x) for x in sys.argv[2].split("x"))
    nt = int(sys.argv[3])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().pl_set_device(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = pl.comm_init_torch()
    assert world == m[0]*m[1]*m[2]
    lat, got, res_d = run(pl, api, size, rank, m, nt)
    coll_ok = collectives(rank, world)
    offs = type("B", (), dict(offsetx=lat.offsetx, offsety=lat.offsety, offsetz=lat.offsetz, nx=lat.nx, ny=lat.ny, nz=lat.nz))
    del lat
    pl.comm_destroy()
    _, want, res_s = run(pl, api, size, 0, (1, 1, 1), nt)
    bad = [k for k in want if not np.array_equal(got[k], block(offs, want[k], size))]
    ok = not bad and abs(res_d - res_s) <= 1e-12*abs(res_s) and coll_ok
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    print(f"rank {rank}: {'OK' if ok else 'MISMATCH ' + ','.join(bad[:8])} residual {res_d:.17g} vs {res_s:.17g}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
