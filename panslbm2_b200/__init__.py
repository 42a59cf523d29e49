"""panslbm2_b200 — B200 (sm_100a) implementation of the PANSLBM2 lattice-Boltzmann forward/adjoint sweep.

The product is `libpanslbm_b200.so` (hand-written CUDA behind the C-ABI of include/panslbm_c.h) plus two host
surfaces over it: the drop-in C++ headers in `panslbm2_b200/src/` (same template surface as the reference) and the
Python mirror in `panslbm2_b200.api`.  There is no CPU fallback; importing works without a GPU, computing does not.
"""
from . import _lib
from ._lib import PanslbmError
from .api import (AAD, AD, ANS, BARRIER, MIRROR, D2Q9, D3Q15, NS, NSin, ConeFilter, DeviceArray, Normalize, Residual, StepPlan, bc_aux, box_sum, collide_args,
                  design_map, gather_field, reduce_absmax, reduce_sum, set_scalar_order,
                  comm_allreduce, comm_destroy, comm_init_loopback, comm_init_torch, halo_describe, synchronize)

__all__ = ["AAD", "AD", "ANS", "BARRIER", "MIRROR", "D2Q9", "D3Q15", "NS", "NSin", "ConeFilter", "DeviceArray", "box_sum", "design_map", "gather_field",
           "reduce_absmax", "reduce_sum", "set_scalar_order", "Normalize", "Residual", "StepPlan", "bc_aux",
           "collide_args", "comm_allreduce", "comm_destroy", "comm_init_loopback", "comm_init_torch", "halo_describe", "synchronize", "PanslbmError"]
