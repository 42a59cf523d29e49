"""Host logic of the checkpoint-recompute state store (panslbm2_b200/transient.py: CheckpointSchedule), simulated without a device:
every state a forward step or an adjoint visit reads must be resident in its slot at that moment, for every (T, every); memory
is ceil-like T/every + every - 1 states; at most one extra forward pass is recomputed."""
import pytest

from panslbm2_b200.transient import CheckpointSchedule


def simulate(T, every, t_hi=None, t_lo=0):
    s = CheckpointSchedule(T, every)
    held = {s.slot(0): 0}          # slot -> the state it holds
    saved = {}                     # checkpoint index -> step the populations were saved after
    pops = 0                       # step whose post-collide populations the forward lattices hold (0 = initial condition)

    def need(t):
        assert held.get(s.slot(t)) == t, f"T={T} every={every}: state {t} is not resident ({s.slot(t)} holds {held.get(s.slot(t))})"

    def step(t):
        nonlocal pops
        assert pops == t - 1, f"T={T} every={every}: step {t} from the populations of step {pops}"
        need(t - 1)
        held[s.slot(t)] = t
        pops = t

    for op, v in s.forward_ops():
        if op == "save":
            assert pops == v*every
            saved[v] = pops
        else:
            step(v)
    assert pops == T
    visited, recomputed = [], 0
    for op, v in s.backward_ops(t_hi, t_lo):
        if op == "restore":
            pops = saved[v]
            need(v*every)
        elif op == "step":
            step(v)
            recomputed += 1
        else:
            need(v)
            if v < T and visited:
                need(v + 1)        # the closures of the visit before read the arrays of step v + 1
            visited.append(v)
    hi = T if t_hi is None else t_hi
    assert visited == list(range(hi, t_lo - 1, -1))
    assert recomputed == s.recomputed_steps(t_hi, t_lo) <= T
    slots = {k for k in held}
    assert len([k for k in slots if k[0] == "perm"]) <= s.n_perm and len([k for k in slots if k[0] == "ring"]) <= s.n_ring
    return s, recomputed


@pytest.mark.parametrize("T", [1, 2, 3, 7, 8, 23, 24, 25, 64, 199])
def test_every_state_is_resident_when_it_is_read(T):
    for every in list(range(1, 12)) + [16, 33, T, T + 1, T + 5]:
        simulate(T, every)
        simulate(T, every, t_hi=T - 1)          # the transient heatsink drivers visit T-1 .. 0 (heatsink3D_transient.cpp:190)
        if T > 3:
            simulate(T, every, t_hi=T - 1, t_lo=2)


def test_memory_and_recompute_bounds():
    s, rec = simulate(199, 16)
    assert s.n_perm + s.n_ring == 199//16 + 1 + 15      # 28 state slots instead of 200
    assert rec == 199 - 199//16 - (199 % 16)            # every ring state outside the last segment once
    s, rec = simulate(199, 1)
    assert s.n_ring == 0 and rec == 0                   # every = 1 is the store-all of the reference
    assert not [op for op in s.forward_ops() if op[0] == "save"]          # ... and needs no population checkpoint
    s = CheckpointSchedule(199, 16)
    assert [v for op, v in s.forward_ops() if op == "save"] == list(range(12))      # the last segment is never recomputed
    s, rec = simulate(50, 400)
    assert s.n_perm == 1 and s.n_ring == 50 and rec == 0
