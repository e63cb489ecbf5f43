"""ctypes binding of libvt_b200.so — the C ABI declared in include/vt_b200.h.

This is the only way Python reaches the CUDA path; there is no CPU fallback.  Loading fails
loudly when the library has not been built (``python -m vlasovtucker_b200.build``).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); must list every symbol of include/vt_b200.h
SIGNATURES = {
    "vt_last_error": (C.c_char_p, []),
    "vt_version": (C.c_int, []),
    "vt_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "vt_ctx_create_group": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "vt_group_size": (C.c_int, [C.c_void_p]),
    "vt_ctx_destroy": (None, [C.c_void_p]),
    "vt_sync": (C.c_int, [C.c_void_p]),
    "vt_device_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "vt_mesh_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_ip, c_dp, c_dp, c_dp, c_ip, c_ip]),
    "vt_species_create": (C.c_int, [C.c_void_p, c_ip, c_dp, c_dp, C.c_double, C.c_double, C.POINTER(C.c_int)]),
    "vt_species_set_params": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double]),
    "vt_species_set_face_bc": (C.c_int, [C.c_void_p, C.c_int, c_u8p, c_u8p, c_ip]),
    "vt_species_set_source_pdfs": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp]),
    "vt_species_set_pdf": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp]),
    "vt_species_get_pdf": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp]),
    "vt_species_set_maxwell": (C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_double, c_dp]),
    "vt_species_set_separable": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp]),
    "vt_measure_dfma_peak": (C.c_int, [C.c_void_p, c_dp]),
    "vt_species_density": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "vt_species_velocity": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "vt_field_set": (C.c_int, [C.c_void_p, c_dp]),
    "vt_field_get": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp]),
    "vt_step_full": (C.c_int, [C.c_void_p, C.c_int, C.c_double, c_dp]),
    "vt_step_full_host": (C.c_int, [C.c_void_p, C.c_int, C.c_double, c_dp, c_dp, c_dp]),
    "vt_step_config": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "vt_step_last_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "vt_launch_count": (C.c_long, [C.c_void_p]),
    "vt_profile_begin": (C.c_int, [C.c_void_p]),
    "vt_profile_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "vt_halo_export": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "vt_halo_attach": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_ip, C.c_void_p]),
    "vt_halo_attach_local": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_ip, C.POINTER(C.c_void_p), c_ip]),
    "vt_halo_set_push": (C.c_int, [C.c_void_p, C.c_int, c_ip, c_ip]),
    "vt_halo_push_current": (C.c_int, [C.c_void_p, C.c_int]),
    "vt_halo_barrier": (C.c_int, [C.c_void_p]),
    "vt_wall_charge_get": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp]),
    "vt_wall_charge_reset": (C.c_int, [C.c_void_p, C.c_int]),
    "vt_poisson_setup": (C.c_int, [C.c_void_p, c_dp, c_dp, c_u8p, c_dp, c_dp]),
    "vt_mesh_set_ghost_geometry": (C.c_int, [C.c_void_p, C.c_int, c_ip, c_ip, c_dp, c_dp, c_dp, c_dp]),
    "vt_poisson_set_global_dirichlet": (C.c_int, [C.c_void_p, C.c_int]),
    "vt_poisson_comm_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vt_poisson_comm_attach": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vt_poisson_comm_attach_local": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vt_poisson_set_push": (C.c_int, [C.c_void_p, c_ip, c_ip]),
    "vt_poisson_update_bc_values": (C.c_int, [C.c_void_p, c_dp, c_dp]),
    "vt_poisson_solve": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp]),
    "vt_poisson_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), c_dp]),
    "vt_charge_density": (C.c_int, [C.c_void_p, c_ip, C.c_int, c_dp]),
    "vt_tucker_enable": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int]),
    "vt_tucker_set_pdf": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "vt_tucker_get_pdf": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "vt_tucker_get_ranks": (C.c_int, [C.c_void_p, C.c_int, c_ip]),
    "vt_tucker_get_factors": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_ip, c_dp, c_dp, c_dp, c_dp]),
    "vt_tucker_density": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "vt_tucker_last_kernel": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "vt_step_tucker": (C.c_int, [C.c_void_p, C.c_int, C.c_double, c_dp]),
    "vt_tucker_halo_export": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "vt_tucker_halo_attach": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vt_tucker_halo_attach_local": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), c_ip]),
}


def lib_path():
    return _build.LIB


def load():
    """Load libvt_b200.so; raises if it is missing (no fallback of any kind)."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build the CUDA extension first (python -m vlasovtucker_b200.build). "
                "vlasovtucker_b200 has no CPU fallback.")
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI and the library disagree
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def u8p(a):
    return None if a is None else a.ctypes.data_as(c_u8p)


def f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


def check(rc):
    if rc:
        raise RuntimeError(load().vt_last_error().decode())
