"""pl_set_scalar_order (include/panslbm_c.h, "Operation order") through the Python mirror: the order is a per-process choice made before
the first lattice exists.  The arithmetic of the scalar order itself is pinned on the CPU (tests/test_hostmath_vs_reference.py against
the reference headers compiled without _USE_AVX_DEFINES) and on the GPU through the drop-in headers (tests/test_gpu_nsopt.py,
tests/test_gpu_dropin.py::test_cpp_surface_built_without_the_avx_macro_...)."""
import pytest

pytestmark = pytest.mark.gpu


def test_scalar_order_is_chosen_before_the_first_lattice():
    import panslbm2_b200 as pl
    from panslbm2_b200 import _lib
    p = pl.D2Q9(8, 6)
    assert _lib.lib().pl_scalar_order() == 0
    with pytest.raises(pl.PanslbmError):
        pl.set_scalar_order(True)
    pl.set_scalar_order(False)        # no change: accepted
    p.free()
