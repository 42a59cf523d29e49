"""Checkpoint-recompute state store for transient forward / adjoint sweeps.

The reference's transient drivers keep one set of macroscopic arrays and one thermal snapshot per time step
(production/heatsink3D_transient.cpp:50-57: rho[t] ... gi[t], 23 doubles per site and step — 194 MB per step at 81 x 161 x 81,
194 GB for nt = 1000) and walk them backwards in the adjoint loop (:190-215).  Keeping every step is the default here too (the
device mirrors of those arrays).  This module is the optional replacement SURVEY.md §8(f)3 asks for: the forward
loop keeps the state of every `every`-th step together with a device checkpoint of the populations (pl_checkpoint_*), the
steps in between go to a ring of `every - 1` state slots that the next segment overwrites; the adjoint loop, walking backwards,
restores the checkpoint in front of a segment whose states are gone and runs the forward plan over it once more.  The
recomputed states are bit-identical to the first ones (same kernels, same inputs), so the adjoint fields and sensitivities are
those of the store-all sweep.  Memory: ceil(T / every) + every - 1 states (+ the population checkpoints) instead of T; time: at
most one extra forward pass.

`CheckpointSchedule` is pure host logic (tests/test_transient_schedule.py runs it without a device); `CheckpointedSweep` executes
it on step plans (panslbm2_b200.api.StepPlan) whose arguments are re-bound every step.
"""
from __future__ import annotations

from . import _lib
from ._lib import check


class CheckpointSchedule:
    """States 0..T: state 0 is the initial condition, state t (1 <= t <= T) is what forward step t stores.  Forward step t reads
    state t - 1 (the closures that follow collide t - 1 read its velocities) and writes state t.  The adjoint visits states
    t_hi, t_hi - 1, ..., t_lo and needs state t and — for the closures of the visit before — state t + 1.

    slot(t) = ("perm", t // every) for t % every == 0, else ("ring", t % every - 1)."""

    def __init__(self, T: int, every: int):
        if T < 1 or every < 1:
            raise ValueError("CheckpointSchedule: T >= 1 and every >= 1")
        self.T, self.every = int(T), int(every)

    @property
    def n_perm(self) -> int:
        return self.T//self.every + 1

    @property
    def n_ring(self) -> int:
        return min(self.every - 1, self.T)

    def slot(self, t: int):
        if not 0 <= t <= self.T:
            raise IndexError(t)
        return ("perm", t//self.every) if t % self.every == 0 else ("ring", t % self.every - 1)

    def segment(self, t: int) -> int:
        return t//self.every

    def forward_ops(self):
        """("save", c): checkpoint the populations in front of segment c (after step c*every; c = 0: after InitialCondition);
        ("step", t): forward step t"""
        last = self.segment(self.T)       # its ring states are still there when the adjoint loop starts: never recomputed
        need = lambda c: self.every >= 2 and c < last
        ops = [("save", 0)] if need(0) else []
        for t in range(1, self.T + 1):
            ops.append(("step", t))
            if t % self.every == 0 and need(t//self.every):
                ops.append(("save", t//self.every))
        return ops

    def backward_ops(self, t_hi: int | None = None, t_lo: int = 0):
        """("restore", c), ("step", t)...: recompute the ring states of segment c; ("visit", t): the adjoint step that uses state t"""
        t_hi = self.T if t_hi is None else t_hi
        ops = []
        valid = self.segment(self.T)      # the segment whose ring states the forward loop left behind
        for t in range(t_hi, t_lo - 1, -1):
            c = self.segment(t)
            if t % self.every != 0 and c != valid:
                ops.append(("restore", c))
                for s in range(c*self.every + 1, min((c + 1)*self.every, self.T + 1)):
                    ops.append(("step", s))
                valid = c
            ops.append(("visit", t))
        return ops

    def recomputed_steps(self, t_hi: int | None = None, t_lo: int = 0) -> int:
        return sum(1 for op, _ in self.backward_ops(t_hi, t_lo) if op == "step")


class Checkpoint:
    """device copy of one lattice's populations with their layout and phase (pl_checkpoint_*)"""

    def __init__(self, lattice):
        self._h = _lib.lib().pl_checkpoint_create(lattice._h)
        if not self._h:
            raise _lib.PanslbmError(_lib.lib().pl_last_error().decode())

    def save(self, lattice):
        check(_lib.lib().pl_checkpoint_save(self._h, lattice._h))
        return self

    def restore(self, lattice):
        check(_lib.lib().pl_checkpoint_restore(self._h, lattice._h))

    def free(self):
        if getattr(self, "_h", None):
            _lib.lib().pl_checkpoint_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CheckpointedSweep:
    """Runs a CheckpointSchedule on a forward StepPlan.

    plan        the finalized forward plan over `lattices` (its collide / closure arrays are re-bound every step)
    make_state  () -> a new state slot (whatever object holds the per-step device arrays; the caller's own type)
    bind        (set_index, state) -> None: plan.rebind(set_index, collide=..., aux=...) with the arrays of `state`
    state0      the slot that holds the initial condition (state 0); made by make_state() if not given

    forward() runs steps 1..T; backward(visit) calls visit(t, state_t, state_t_plus_1) for t = t_hi..t_lo with both states
    resident (state_t_plus_1 is None for t = T), recomputing segments as it goes."""

    def __init__(self, plan, lattices, T, every, make_state, bind, state0=None, checkpoint=None):
        self.plan, self.lattices, self.bind = plan, list(lattices), bind
        Checkpoint_ = checkpoint if checkpoint is not None else Checkpoint      # (a test double stands in for the device copy on a CPU)
        self.sched = CheckpointSchedule(T, every)
        self.perm = [state0 if (c == 0 and state0 is not None) else make_state() for c in range(self.sched.n_perm)]
        self.ring = [make_state() for _ in range(self.sched.n_ring)]
        # every population checkpoint the schedule will take, allocated up front (no device allocation inside the loops)
        self.cps = {v: ([Checkpoint_(l) for l in self.lattices], 0) for op, v in self.sched.forward_ops() if op == "save"}
        self.recomputed = 0

    def state(self, t):
        kind, i = self.sched.slot(t)
        return self.perm[i] if kind == "perm" else self.ring[i]

    def _step(self, t, last=False):
        self.bind(self.plan.next_set(), self.state(t))
        self.plan.advance(1, end_streamed=last)

    def forward(self, end_streamed=True):
        for op, v in self.sched.forward_ops():
            if op == "save":
                cps = self.cps[v][0]
                for cp, l in zip(cps, self.lattices):
                    cp.save(l)
                self.cps[v] = (cps, self.plan.parity)
            else:
                self._step(v, last=(end_streamed and v == self.sched.T))

    def backward(self, visit, t_hi=None, t_lo=0):
        T = self.sched.T
        for op, v in self.sched.backward_ops(t_hi, t_lo):
            if op == "restore":
                cps, parity = self.cps[v]
                for cp, l in zip(cps, self.lattices):
                    cp.restore(l)
                self.plan.set_parity(parity)
                # the closures that follow collide c*every read the arrays of that step: the set of the last collide
                self.bind(parity, self.state(v*self.sched.every))
            elif op == "step":
                self._step(v)
                self.recomputed += 1
            else:
                visit(v, self.state(v), self.state(v + 1) if v < T else None)

    def free(self):
        for cps, _ in self.cps.values():
            for cp in cps:
                cp.free()
        self.cps = {}
