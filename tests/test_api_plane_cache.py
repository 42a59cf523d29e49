"""The Python mirror bakes a boundary plane from callables once per (plane, callable VALUE): same code object, same captured
scalars / scalar dicts / nested functions, same defaults and global scalars.  Anything whose value cannot be pinned down (arrays,
objects) makes the callable uncacheable, so that a changed capture can never be served a stale plane."""
import numpy as np

from panslbm2_b200.api import _callable_signature as sig


def make(L, qn, extra=None):
    inL = lambda i, k: (i < L) & (k < L)
    p = {"qn0": qn, "tem0": 0.0}
    if extra is not None:
        return lambda i, j, k: extra[0] + 0*i
    return lambda i, j, k: np.where((j == 0) & inL(i, k), p["qn0"], 0.0)


def test_same_code_same_captures_is_one_plane():
    assert sig(make(3.0, 1e-2)) is not None
    assert sig(make(3.0, 1e-2)) == sig(make(3.0, 1e-2))


def test_changed_capture_changes_the_signature():
    assert sig(make(3.0, 1e-2)) != sig(make(4.0, 1e-2))      # nested function's capture
    assert sig(make(3.0, 1e-2)) != sig(make(3.0, 2e-2))      # scalar inside a captured dict


def test_unpinnable_captures_are_not_cached():
    assert sig(make(3.0, 1e-2, extra=np.zeros(3))) is None
    assert sig(None) == ("none",)
    assert sig(np.add) is None      # no code object
