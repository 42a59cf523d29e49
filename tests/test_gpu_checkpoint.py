"""Checkpoint-recompute for transient adjoints (include/panslbm_c.h pl_checkpoint_*, panslbm2_b200/transient.py).
  * pl_checkpoint_save / restore: a plan continued from a restored checkpoint repeats the steps it made the first time, bit for
    bit, whichever layout and phase the populations were saved in;
  * the transient heatsink loops of BASELINE configs[4] (production/heatsink3D_transient.cpp:145-232) with every step stored
    (every = 1, what the reference does) and with every K-th step stored and the rest recomputed during the adjoint loop: both
    reproduce the fixtures generated from the reference headers (tests/golden/transient.npz) bit for bit."""
import hashlib
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


@pytest.mark.parametrize("first", [0, 1, 4, 5])
def test_restored_checkpoint_repeats_the_same_steps(first):
    import panslbm2_b200 as pl
    from panslbm2_b200 import api
    from panslbm2_b200.transient import Checkpoint
    lx, ly, lz, more = 18, 13, 11, 7
    nu, u0 = 0.1, 0.1
    wall = lambda i, j, k: np.where((i == 0) | (i == lx - 1) | (j == 0) | (j == ly - 1) | (k == 0), 1, 0)
    lid = lambda i, j, k: k == lz - 1
    uvals = [lambda i, j, k: 0.0*u0, lambda i, j, k: u0, lambda i, j, k: 0.0]
    pf = pl.D3Q15(lx, ly, lz)
    N = pf.nxyz
    rho = pl.DeviceArray(N, 1.0)
    u = [pl.DeviceArray(N, 0.0) for _ in range(3)]
    pl.NS.InitialCondition(pf, rho, *u)
    plan = pl.StepPlan(pf).set_collide(pl.collide_args(api.M_NS_COLLIDE, True, nu, rho=rho, ux=u[0], uy=u[1], uz=u[2]))
    plan.add_bounce(pf, wall).add_closure(pf, api.BC_NS_SET_U, lid, uvals).set_smooth_corner(True).finalize()
    if first:
        plan.advance(first, end_streamed=False)      # odd / even counts leave the one buffer in the streamed / natural layout
    cp = Checkpoint(pf).save(pf)
    parity = plan.parity
    plan.advance(more, end_streamed=True)
    a = [rho.to_host()] + [x.to_host() for x in u] + list(pf.get_populations())
    # trash the lattice and the outputs, then go back
    pl.NS.InitialCondition(pf, rho, *u)
    for x in [rho] + u:
        x.fill(-3.0)
    cp.restore(pf)
    plan.set_parity(parity)
    plan.advance(more, end_streamed=True)
    b = [rho.to_host()] + [x.to_host() for x in u] + list(pf.get_populations())
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert np.max(np.abs(a[2])) > 1e-3
    cp.free()
    other = pl.D3Q15(lx + 1, ly, lz)
    with pytest.raises(pl.PanslbmError):
        Checkpoint(pf).restore(pf)                   # nothing saved yet
    c2 = Checkpoint(pf).save(pf)
    with pytest.raises(pl.PanslbmError):
        c2.restore(other)                            # another shape


def cases():
    spec = importlib.util.spec_from_file_location("make_transient_golden", os.path.join(G, "make_transient_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.TRANSIENT_CASES


def check(tag, res):
    z = np.load(os.path.join(G, "transient.npz"))
    keys = sorted(k.split("/")[1] for k in z.files if k.startswith(tag + "/") and k.endswith("/sha"))
    assert len(keys) >= 25
    for k in keys:
        a = res[k] + 0.0
        if k == "extra":        # the objective: a sum over the heat patch of every step, reduced on the device (other summation order)
            want = z[f"{tag}/extra/s5"]
            assert abs(a[0] - want[0]) <= 1e-12*abs(want[0]), (a, want)
            continue
        assert np.array_equal(a[::5], z[f"{tag}/{k}/s5"]), f"{tag}: {k} differs from the reference fixture (max abs {np.max(np.abs(a[::5] - z[f'{tag}/{k}/s5'])):.3e})"
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest() == bytes(z[f"{tag}/{k}/sha"]), f"{tag}: {k} digest"


@pytest.mark.parametrize("tag,every", [("tr3d", 1), ("tr3d", 4), ("tr3d", 7), ("tr3d", 23), ("tr3d_tail", 3), ("tr3d_tail", 5)])
def test_checkpointed_transient_sweep_matches_reference_fixture(tag, every):
    from panslbm2_b200.transient import CheckpointSchedule
    from transient_case import run_transient_cuda
    dim, size, nt = cases()[tag]
    assert dim == 3
    stats = {}
    res = run_transient_cuda(size, nt, every, stats)
    check(tag, res)
    s = CheckpointSchedule(nt - 1, every)
    assert stats["recomputed"] == s.recomputed_steps(nt - 2, 0)
    assert stats["states"] == s.n_perm + s.n_ring
    if every in (4, 7, 3, 5):
        assert stats["recomputed"] > 0 and stats["states"] < nt - 1
