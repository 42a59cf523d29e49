"""GPU parity of the pieces around the sweep that an optimisation iteration needs on the device:
the cone filters through the Python mirror (weight patterns) against the fixtures of the reference's serial host filters, the design
map of production/heatsink3D.cpp:114-119, the patch objective (:227-240) and the plain reductions."""
import os

import numpy as np
import pytest

import filter_case as FC
import heatsink_case as H
from helpers import gcoords

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("tag", ["hs3d_box", "hs2d_box", "cone3d", "cone2d_r3"])
def test_cone_filter_patterns_match_reference_filters(tag):
    import panslbm2_b200 as pl
    dim, size, R, beta, box = FC.CASES[tag]
    v, d = FC.inputs(tag)
    p = pl.D3Q15(*size) if dim == 3 else pl.D2Q9(size[0], size[1])
    weight = None
    if box:
        bx, by, bz = box

        def weight(i1, j1, k1, i2, j2, k2):      # production/heatsink3D.cpp:87-93
            inside = (i1 < bx) & (j1 < by) & (k1 < bz) & (i2 < bx) & (j2 < by) & (k2 < bz)
            cone = (R - np.sqrt((i1 - i2)**2.0 + (j1 - j2)**2.0 + (k1 - k2)**2.0))/R
            return np.where(inside, cone, np.where((i1 == i2) & (j1 == j2) & (k1 == k2), 1.0, 0.0))
    f = pl.ConeFilter(p, R, weight)
    n = size[0]*size[1]*size[2]
    assert f.npatterns < n      # the per-site table is gone: a handful of patterns
    dv, dd = pl.DeviceArray.from_host(v), pl.DeviceArray.from_host(d)
    z = np.load(os.path.join(G, "filters.npz"))
    assert np.array_equal(f.density(dv).to_host(), z[f"{tag}/fv"])                     # no transcendental: bit-exact
    assert np.max(np.abs(f.heaviside(dv, beta).to_host() - z[f"{tag}/rho"])) <= 1e-14
    assert np.max(np.abs(f.heaviside_sensitivity(dv, dd, beta).to_host() - z[f"{tag}/dfds"])) <= 1e-13*np.max(np.abs(z[f"{tag}/dfds"]))


def test_filter_pattern_count_does_not_grow_with_the_lattice():
    import panslbm2_b200 as pl
    R = 2.4
    counts = []
    for size in ((20, 18, 16), (40, 36, 32)):
        bx, by, bz = [3*(s - 1)//4 + 1 for s in size]

        def weight(i1, j1, k1, i2, j2, k2):
            inside = (i1 < bx) & (j1 < by) & (k1 < bz) & (i2 < bx) & (j2 < by) & (k2 < bz)
            cone = (R - np.sqrt((i1 - i2)**2.0 + (j1 - j2)**2.0 + (k1 - k2)**2.0))/R
            return np.where(inside, cone, np.where((i1 == i2) & (j1 == j2) & (k1 == k2), 1.0, 0.0))
        counts.append(pl.ConeFilter(pl.D3Q15(*size), R, weight).npatterns)
    assert counts[0] == counts[1] and counts[0] <= 5**3 + 2, counts


def test_design_map_is_bit_identical_to_the_drivers_formulas():
    import panslbm2_b200 as pl
    size = (17, 13, 11)
    p = H.params(3, size)
    i, j, k = gcoords(*size)
    inbox = (i < p["mx"]) & (j < p["my"]) & (k < p["mz"])
    ss = np.where(inbox, 0.5 + 0.4*np.sin(0.37*i)*np.cos(0.23*j)*np.sin(0.31*k + 0.5), 1.0)
    want = H.design_fields(p, i, j, k)      # alpha, kappa, dads, dkds
    kappa, alpha, dkds, dads = pl.design_map(pl.DeviceArray.from_host(ss), p["diff_fluid"], p["diff_solid"], p["qg"], p["alphamax"]/float(p["ly"] - 1), p["qf"])
    for name, got, w in (("alpha", alpha, want[0]), ("kappa", kappa, want[1]), ("dads", dads, want[2]), ("dkds", dkds, want[3])):
        assert np.array_equal(got.to_host(), w), name


def test_reductions_and_patch_objective():
    import panslbm2_b200 as pl
    size = (19, 14, 12)
    n = size[0]*size[1]*size[2]
    rs = np.random.RandomState(7)
    v = rs.uniform(-1.0, 2.0, n)
    d = pl.DeviceArray.from_host(v)
    assert abs(pl.reduce_sum(d) - v.sum()) <= 1e-13*np.abs(v).sum()
    assert pl.reduce_absmax(d) == np.abs(v).max()
    p = pl.D3Q15(*size)
    L = 5
    want = v.reshape(size[2], size[1], size[0])[:L, 0, :L].sum()       # heatsink3D.cpp:229-235: i < L, j == 0, k < L
    got = pl.box_sum(p, d, 0, L, 0, 1, 0, L)
    assert abs(got - want) <= 1e-13*max(1.0, abs(want))
    assert pl.box_sum(p, d, 4, 4, 0, 1, 0, L) == 0.0
    assert np.array_equal(pl.gather_field(p, d), v)                    # one block: the field itself
