"""One population buffer per lattice (north_star: "one fused stream+collide pass per time step in an AA-pattern ... in-place layout";
the reference keeps two, f and the hidden fnext, d3q15.h:41-45, 238).  The fused passes of a plan alternate between a gather pass
(natural layout -> streamed layout) and a local pass (back), every thread writing exactly the locations it read.  Parity of the
passes themselves is what every fused-plan test of the suite checks (they all run in place); here: the memory really is one buffer,
every other operation still sees the natural layout at any point of the loop, and the two-buffer schedule gives the same numbers."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import heatsink_case as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mem():
    from panslbm2_b200 import _lib
    out = (C.c_uint64*4)()
    _lib.check(_lib.lib().pl_memory_stats(out))
    return [int(v) for v in out]


def test_steady_loop_holds_one_buffer_per_lattice():
    import bench
    import panslbm2_b200 as pl
    from panslbm2_b200 import _lib, api
    import gc
    gc.collect()
    _lib.lib().pl_memory_trim()
    base = mem()
    size = (40, 36, 32)
    sw = bench.HeatsinkSweep(pl, api, size)
    sw.upload_design()
    sw.init_forward()
    per_lattice = 15*8*((sw.n + 15)//16*16)
    assert mem()[0] - base[0] == 2*per_lattice
    sw.f.Stream(); sw.g.Stream()                                          # a standalone Stream() borrows a spare (f and g share it) ...
    assert mem()[1] == per_lattice
    sw.init_forward()
    sw.fplan.advance(71, end_streamed=False, save_last=2)
    m = mem()
    assert m[0] - base[0] == 2*per_lattice and m[1] == 0, m          # ... which a steady loop gives back: no second buffer anywhere
    borrows = m[2]
    sw.fplan.advance(7, end_streamed=False, save_last=2)
    assert mem()[2] == borrows                                           # ... and none was borrowed on the way
    # observing the populations in the middle of the loop (77 passes so far: streamed layout inside) converts through ONE spare
    conv = mem()[3]
    f0a, fa = sw.f.get_populations()
    m = mem()
    assert m[3] == conv + 1 and m[1] <= per_lattice, m
    sw.fplan.advance(4, end_streamed=True, save_last=2)
    f0b, fb = sw.f.get_populations()
    assert np.isfinite(fb).all() and not np.array_equal(fa, fb)


@pytest.mark.parametrize("nt", [6, 7])
def test_observation_at_any_pass_parity_sees_the_natural_layout(nt):
    """stop after an even / odd number of fused passes (natural / streamed layout inside), look at the populations, go on: the
    same as never having looked, and the same as the call-by-call loop"""
    size = (13, 11, 10)
    a = H.run_cuda(3, size, nt + 4, fused=True, only_forward=True, chunks=(nt + 4, 1))
    import panslbm2_b200 as pl

    looks = []

    def look(A, gsnap, igsnap):
        looks.append(1)
    b = H.run_cuda(3, size, nt + 4, fused=True, only_forward=True, chunks=(nt, 4), observe=look)
    H.compare(b, a, "chunks")
    c = H.run_cuda(3, size, nt + 4, fused=False, only_forward=True)
    H.compare(b, c, "fused vs call by call")


@pytest.mark.parametrize("dim,size,nt,chunks", [(2, (33, 27, 1), 41, (7, 9)), (3, (15, 13, 11), 23, (5, 8))])
def test_cooperative_multi_step_launch_gives_the_same_numbers(dim, size, nt, chunks):
    """PANSLBM_COOP_SITES=n runs the fused passes of one advance call in ONE cooperative launch on lattices of up to n sites (k_steps:
    grid barriers instead of kernel boundaries; an experiment that measured slower and is off by default) — same numbers as pass by
    pass, and both equal the call-by-call loop"""
    code = ("import sys, json, hashlib, numpy as np\n"
            f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
            "import heatsink_case as H\n"
            "from panslbm2_b200 import _lib\n"
            f"r = H.run_cuda({dim}, {size!r}, {nt}, fused=FUSED, chunks={chunks!r}, save_last=SAVE)\n"
            "d = {k: hashlib.sha256(np.ascontiguousarray(v + 0.0).tobytes()).hexdigest() for k, v in sorted(r.items())}\n"
            "d['launches'] = int(_lib.lib().pl_launch_count())\n"
            "print(json.dumps(d))\n")
    import json
    outs = {}
    for tag, env, fused, save in (("coop", {"PANSLBM_COOP_SITES": "400000"}, True, 2), ("passes", {"PANSLBM_COOP_SITES": "0"}, True, 2), ("calls", {}, False, None)):
        r = subprocess.run([sys.executable, "-c", code.replace("FUSED", str(fused)).replace("SAVE", str(save))], capture_output=True, text=True,
                           env=dict(os.environ, **env), timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = json.loads(r.stdout.strip().splitlines()[-1])
    launches = {k: v.pop("launches") for k, v in outs.items()}
    assert outs["coop"] == outs["passes"] == outs["calls"]
    assert launches["coop"] < launches["passes"], launches      # the batches really went through k_steps


def test_two_buffer_schedule_gives_the_same_numbers():
    code = ("import sys, json, hashlib, numpy as np\n"
            f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
            "import heatsink_case as H\n"
            "r = H.run_cuda(3, (14, 12, 11), 9, fused=True, chunks=(2, 3), save_last=2)\n"
            "print(json.dumps({k: hashlib.sha256(np.ascontiguousarray(v + 0.0).tobytes()).hexdigest() for k, v in sorted(r.items())}))\n")
    outs = []
    for knob in ("1", "0"):
        env = dict(os.environ, PANSLBM_INPLACE=knob)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1]
