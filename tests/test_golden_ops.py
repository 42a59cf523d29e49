"""Committed fixtures generated from the reference build (tests/golden/make_golden.py) against the product's site math
compiled for the host (tests/hostmath).  Runs anywhere — no /root/reference, no oracle/_ref, no GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

import scenarios as S
from helpers import hostmath_backend

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DIGESTS = json.load(open(os.path.join(G, "ops_digests.json")))


def digest(res):
    h = hashlib.sha256()
    for _, a in res:
        h.update(np.ascontiguousarray(a + 0.0, dtype=np.float64).tobytes())
    return h.hexdigest()


def run(be, key):
    tag, what, *rest = key.split("/")
    dim = int(tag[1])
    size = tuple(int(v) for v in tag.split("_")[1].split("x"))
    if what == "collide":
        return S.collide(be, dim, rest[0], size, 3)
    if what == "closure":
        return S.closure(be, dim, rest[0], size, 5)
    if what == "sensitivity":
        return S.sensitivity(be, dim, rest[0], size, 9)
    return S.inits(be, dim, size, 2)


@pytest.mark.parametrize("key", sorted(DIGESTS))
def test_hostmath_matches_reference_fixture(key):
    be = hostmath_backend(int(key[1]))
    assert digest(run(be, key)) == DIGESTS[key], key
