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


# ---------------------------------------------------------------------------------------------------------
# The executor (CheckpointedSweep) on a toy model of the plan semantics, no device: a "lattice" is a number, a step of the plan is
#   pops <- F(pops, closure arrays of the set of the last collide) ; state[t] <- G(pops)      (the fused pass of panslbm_api.cu:
#   Stream + closures with the arguments of step t - 1, then the collide of step t with the arguments of step t)
# so a wrong binding, a stale ring slot or a checkpoint restored with the wrong argument-set index changes the numbers.
class ToyLattice:
    def __init__(self):
        self.pops, self.streamed = 1.0, 1


class ToyCheckpoint:
    def __init__(self, lattice):
        self.saved = None

    def save(self, lattice):
        self.saved = (lattice.pops, lattice.streamed)
        return self

    def restore(self, lattice):
        lattice.pops, lattice.streamed = self.saved

    def free(self):
        pass


class ToyPlan:
    """two argument sets; `parity` = set of the last collide after a collide, of the next one in the streamed phase"""

    def __init__(self, lattice):
        self.l, self.sets, self.parity, self.steps = lattice, [None, None], 0, 0

    def next_set(self):
        return self.parity if self.l.streamed else self.parity ^ 1

    def set_parity(self, p):
        self.parity = p

    def bind(self, k, state):
        self.sets[k] = state

    def advance(self, n, end_streamed=False):
        assert n == 1
        if self.l.streamed:                         # first collide after InitialCondition: no Stream in front of it
            k = self.parity
        else:
            prev = self.sets[self.parity]           # closures read the velocities the last collide stored
            self.l.pops = 0.75*self.l.pops + 0.125*prev["u"] + 0.01
            k = self.parity ^ 1
            self.parity = k
        st = self.sets[k]
        st["u"] = 0.5*self.l.pops + 0.001*self.steps_in(st)      # what the collide stores
        st["t"] = st["want_t"]
        self.l.pops = 0.9*self.l.pops + 0.05
        self.l.streamed = 0
        self.steps += 1
        if end_streamed:
            self.l.pops = 0.75*self.l.pops + 0.125*st["u"] + 0.01
            self.l.streamed = 1
            self.parity ^= 1

    @staticmethod
    def steps_in(st):
        return st["want_t"]


def run_toy(T, every):
    from panslbm2_b200.transient import CheckpointedSweep
    lat = ToyLattice()
    plan = ToyPlan(lat)
    s0 = {"u": 0.0, "t": 0, "want_t": 0}
    current = {"t": 0}

    def make_state():
        return {"u": None, "t": None, "want_t": None}

    def bind(k, state):
        plan.bind(k, state)
    sweep = CheckpointedSweep(plan, [lat], T, every, make_state, bind, state0=s0, checkpoint=ToyCheckpoint)
    # the executor binds slot objects; tell each slot which step it is about to hold (the real bind() passes the step's arrays)
    orig_step = sweep._step

    def step(t, last=False):
        sweep.state(t)["want_t"] = t
        orig_step(t, last)
    sweep._step = step
    sweep.forward(end_streamed=False)
    seen = []

    def visit(t, st, st_next):
        assert st["t"] == t, (T, every, t, st)
        if st_next is not None:
            assert st_next["t"] == t + 1, (T, every, t, st_next)
        seen.append((t, st["u"]))
    sweep.backward(visit, t_hi=T - 1, t_lo=0)
    return seen, sweep.recomputed


@pytest.mark.parametrize("T", [1, 2, 5, 12, 23, 40])
def test_executor_reproduces_the_store_all_sweep_on_a_toy_plan(T):
    want, rec = run_toy(T, 1)
    assert rec == 0 and [t for t, _ in want] == list(range(T - 1, -1, -1))
    for every in (2, 3, 4, 7, 8, T, T + 3):
        got, rec = run_toy(T, every)
        assert got == want, (T, every)          # every visited state carries exactly the value the store-all sweep saw (bit for bit)
        assert rec == CheckpointSchedule(T, every).recomputed_steps(T - 1, 0)
