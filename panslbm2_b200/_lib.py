"""ctypes binding of the C-ABI in include/panslbm_c.h.  Loading fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PANSLBM_LIB_TAG=<tag>: an A/B build variant (build.py, PANSLBM_BUILD_TAG) instead of the library; tuning experiments only
_TAG = os.environ.get("PANSLBM_LIB_TAG", "")
LIB_PATH = os.path.join(HERE, "libpanslbm_b200" + ("_" + _TAG if _TAG else "") + ".so")

c_double_p = C.POINTER(C.c_double)


class CollideArgs(C.Structure):
    _fields_ = ([("model", C.c_int), ("issave", C.c_int), ("viscosity", C.c_double), ("diffusivity_const", C.c_double),
                 ("gx", C.c_double), ("gy", C.c_double), ("gz", C.c_double), ("tem0", C.c_double)]
                + [(n, C.c_void_p) for n in ("alpha", "diffusivity", "beta", "dirx", "diry", "dirz",
                                             "rho", "ux", "uy", "uz", "tem", "qx", "qy", "qz",
                                             "ip", "iux", "iuy", "iuz", "imx", "imy", "imz", "item", "iqx", "iqy", "iqz",
                                             "snapshot")])


class BcAux(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rho", "ux", "uy", "uz", "tem", "diffusivity")] + [("diffusivity_const", C.c_double), ("eps", C.c_double)]


class SensArgs(C.Structure):
    _fields_ = [("kind", C.c_int), ("dfds", C.c_void_p)] + [(n, C.c_void_p) for n in (
        "ux", "uy", "uz", "imx", "imy", "imz", "dads", "tem", "item", "iqx", "iqy", "iqz", "gsnap", "igsnap", "diffusivity", "dkds", "dbds")]


_PROTOS = {
    "pl_last_error": (C.c_char_p, []),
    "pl_version": (C.c_char_p, []),
    "pl_device_count": (C.c_int, []),
    "pl_set_device": (C.c_int, [C.c_int]),
    "pl_synchronize": (C.c_int, []),
    "pl_get_stream": (C.c_void_p, []),
    "pl_set_stream": (C.c_int, [C.c_void_p]),
    "pl_launch_count": (C.c_uint64, []),
    "pl_launch_count_reset": (None, []),
    "pl_comm_unique_id": (C.c_int, [C.c_char_p]),
    "pl_comm_init": (C.c_int, [C.c_char_p, C.c_int, C.c_int]),
    "pl_comm_init_loopback": (C.c_int, [C.c_int]),
    "pl_comm_destroy": (C.c_int, []),
    "pl_comm_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "pl_comm_allreduce": (C.c_int, [c_double_p, C.c_int, C.c_int]),
    "pl_comm_allreduce_v": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "pl_comm_p2p": (C.c_int, [C.c_void_p, C.c_int]),
    "pl_halo_describe": (C.c_int, [C.c_int] * 9 + [C.c_void_p, C.POINTER(C.c_int)]),
    "pl_array_alloc": (C.c_void_p, [C.c_size_t]),
    "pl_array_free": (C.c_int, [C.c_void_p]),
    "pl_array_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "pl_array_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "pl_array_upload_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "pl_array_download_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "pl_copy_fence": (C.c_int, []),
    "pl_copy_wait": (C.c_int, []),
    "pl_array_fill": (C.c_int, [C.c_void_p, C.c_double, C.c_size_t]),
    "pl_lattice_create": (C.c_void_p, [C.c_int] * 8),
    "pl_lattice_destroy": (C.c_int, [C.c_void_p]),
    "pl_lattice_info": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pl_lattice_set_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pl_lattice_get_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pl_lattice_device_view": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "pl_lattice_streamed": (C.c_int, [C.c_void_p]),
    "pl_set_scalar_order": (C.c_int, [C.c_int]),
    "pl_scalar_order": (C.c_int, []),
    "pl_checkpoint_create": (C.c_void_p, [C.c_void_p]),
    "pl_checkpoint_save": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pl_checkpoint_restore": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pl_checkpoint_destroy": (C.c_int, [C.c_void_p]),
    "pl_memory_stats": (C.c_int, [C.POINTER(C.c_uint64)]),
    "pl_memory_trim": (C.c_int, []),
    "pl_in_call": (C.c_int, []),
    "pl_stream": (C.c_int, [C.c_void_p, C.c_int]),
    "pl_smooth_corner": (C.c_int, [C.c_void_p]),
    "pl_smooth_corner_at": (C.c_int, [C.c_void_p] + [C.c_int] * 6),
    "pl_bc_create": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pl_bc_destroy": (C.c_int, [C.c_void_p]),
    "pl_bc_is_empty": (C.c_int, [C.c_void_p]),
    "pl_bc_update_values": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pl_bc_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(BcAux)]),
    "pl_collide": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(CollideArgs)]),
    "pl_snapshot_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pl_snapshot_from_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pl_snapshot_convert": (C.c_int, [C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int]),
    "pl_initial_condition": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int]),
    "pl_plan_create": (C.c_void_p, [C.c_void_p, C.c_void_p]),
    "pl_plan_destroy": (C.c_int, [C.c_void_p]),
    "pl_plan_set_collide": (C.c_int, [C.c_void_p, C.POINTER(CollideArgs), C.POINTER(CollideArgs)]),
    "pl_plan_set_stream": (C.c_int, [C.c_void_p, C.c_int]),
    "pl_plan_add_bc": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(BcAux), C.POINTER(BcAux)]),
    "pl_plan_set_smooth_corner": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "pl_plan_add_smooth_corner_at": (C.c_int, [C.c_void_p] + [C.c_int]*7),
    "pl_plan_finalize": (C.c_int, [C.c_void_p]),
    "pl_plan_advance": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "pl_plan_advance_observed": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "pl_plan_rebind": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(CollideArgs), C.POINTER(BcAux), C.c_int]),
    "pl_plan_parity": (C.c_int, [C.c_void_p]),
    "pl_plan_set_parity": (C.c_int, [C.c_void_p, C.c_int]),
    "pl_plan_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "pl_plan_profile_read": (C.c_int, [C.c_void_p, c_double_p, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "pl_plan_profile_read2": (C.c_int, [C.c_void_p, c_double_p, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "pl_residual": (C.c_int, [C.c_void_p] * 6 + [C.c_size_t, c_double_p]),
    "pl_reduce_sum": (C.c_int, [C.c_void_p, C.c_size_t, c_double_p]),
    "pl_reduce_absmax": (C.c_int, [C.c_void_p, C.c_size_t, c_double_p]),
    "pl_normalize": (C.c_int, [C.c_void_p, C.c_size_t]),
    "pl_sensitivity": (C.c_int, [C.c_void_p, C.POINTER(SensArgs)]),
    "pl_sensitivity_heat_source": (C.c_int, [C.c_void_p] * 9),
    "pl_filter_create": (C.c_void_p, [C.c_void_p, C.c_int, C.c_void_p]),
    "pl_filter_create_patterns": (C.c_void_p, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "pl_filter_patterns": (C.c_int, [C.c_void_p]),
    "pl_reduce_box_sum": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int]*6 + [c_double_p]),
    "pl_design_map": (C.c_int, [C.c_void_p, C.c_size_t] + [C.c_double]*5 + [C.c_void_p]*4),
    "pl_comm_gather_field": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pl_filter_destroy": (C.c_int, [C.c_void_p]),
    "pl_filter_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    # host-pointer surface (bound by the C++ drop-in headers; declared here so that the export test covers it)
    "plh_last_error": (C.c_char_p, []),
    "plh_alloc": (C.c_void_p, [C.c_size_t]),
    "plh_free": (None, [C.c_void_p]),
    "plh_owns": (C.c_int, [C.c_void_p]),
    "plh_owns_range": (C.c_int, [C.c_void_p]),
    "plh_bc_update_values": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "plh_lattice_attach_views": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "plh_lattice_detach": (C.c_int, [C.c_void_p]),
    "plh_collide": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(CollideArgs)]),
    "plh_stream": (C.c_int, [C.c_void_p, C.c_int]),
    "plh_smooth_corner": (C.c_int, [C.c_void_p]),
    "plh_smooth_corner_at": (C.c_int, [C.c_void_p] + [C.c_int] * 6),
    "plh_bc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(BcAux)]),
    "plh_initial_condition": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int]),
    "plh_residual": (C.c_int, [C.c_void_p] * 6 + [C.c_size_t, c_double_p]),
    "plh_normalize": (C.c_int, [C.c_void_p, C.c_size_t]),
    "plh_sensitivity": (C.c_int, [C.c_void_p, C.POINTER(SensArgs)]),
    "plh_sensitivity_heat_source": (C.c_int, [C.c_void_p] * 9),
    "plh_filter_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "plh_host_acquire": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int]),
    "plh_sync": (C.c_int, []),
    "plh_stats": (C.c_int, [C.POINTER(C.c_uint64)]),
    "plh_store_stats": (C.c_int, [C.POINTER(C.c_uint64)]),
}
EXPORTS = tuple(_PROTOS)

_lib = None


class PanslbmError(RuntimeError):
    pass


def lib():
    """The loaded library with prototypes set.  Raises if it has not been built (python -m panslbm2_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PanslbmError(f"{LIB_PATH} is missing: build it with `python -m panslbm2_b200.build` (nvcc, sm_100a). "
                               "There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(code: int):
    if code != 0:
        raise PanslbmError(lib().pl_last_error().decode())


def require_device():
    if lib().pl_device_count() < 1:
        raise PanslbmError("no CUDA device visible: panslbm2_b200 has no CPU path")
