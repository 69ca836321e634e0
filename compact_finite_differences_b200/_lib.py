"""ctypes binding of libcfd_b200.so (include/cfd_b200.h).  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcfd_b200.so")

CFD_OK, CFD_EINVAL, CFD_ECUDA, CFD_EUNSUPPORTED, CFD_ETIMEOUT = 0, -1, -2, -3, -4
CFD_IPC_HANDLE_BYTES = 64

# every symbol include/cfd_b200.h declares: (restype, argtypes)
_i, _l, _d, _vp = ctypes.c_int, ctypes.c_long, ctypes.c_double, ctypes.c_void_p
_dp = ctypes.POINTER(ctypes.c_double)
_pp = ctypes.POINTER(ctypes.c_void_p)
SIGNATURES = {
    "cfd_version": (_i, []),
    "cfd_last_error": (ctypes.c_char_p, []),
    "cfd_create": (_i, [_pp, _i, _i, _i, _i, _d, _i, _i]),
    "cfd_destroy": (None, [_vp]),
    "cfd_create_scheme": (_i, [_pp, _i, _i, _i, _i, _d, _i]),
    "cfd_plan_lookahead": (_i, [_vp]),
    "cfd_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "cfd_apply_xy": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "cfd_apply_xyz": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cfd_plan_set_xy_warps": (_i, [_vp, _i]),
    "cfd_compute_rhs": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "cfd_plan_coeffs": (_i, [_vp, _dp]),
    "cfd_sum_solutions": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "cfd_interface_pack": (_i, [_vp, _vp, _vp, _vp]),
    "cfd_reduced_correct": (_i, [_vp, _vp, _vp, _vp]),
    "cfd_edge_faces": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "cfd_apply_coupled": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cfd_reduced_unknowns": (_i, [_vp, _vp, _i, _vp, _vp, _vp, ctypes.c_ulonglong, _vp]),
    "cfd_nb_layout": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "cfd_debug_neighbour": (_i, [_i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), _dp, _dp, _dp]),
    "cfd_push_planes": (_i, [_vp, _vp, _vp, _vp, _l, _vp, _vp, ctypes.c_ulonglong, _vp]),
    "cfd_edge_faces_p2p": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_ulonglong, _vp]),
    "cfd_wait_flags": (_i, [_vp, _vp, ctypes.c_ulonglong, _vp]),
    "cfd_edge_faces_push": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_ulonglong, _vp]),
    "cfd_reduced_unknowns_deferred": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_ulonglong, _vp]),
    "cfd_apply_host": (_i, [_vp, _vp, _vp, _i]),
    "cfd_pthomas_create": (_i, [_pp, _dp, _dp, _dp, _i]),
    "cfd_pthomas_solve": (_i, [_vp, _vp, _l, _vp]),
    "cfd_pthomas_destroy": (None, [_vp]),
    "cfd_zpart_create": (_i, [_pp, _vp]),
    "cfd_zpart_export": (_i, [_vp, _vp]),
    "cfd_zpart_connect": (_i, [_vp, _vp, _vp]),
    "cfd_zpart_buffer": (_vp, [_vp]),
    "cfd_zpart_connect_ptr": (_i, [_vp, _vp, _vp]),
    "cfd_zpart_set_ctas": (_i, [_vp, _i]),
    "cfd_zpart_begin": (_i, [_vp, _vp, _vp]),
    "cfd_zpart_apply": (_i, [_vp, _vp, _vp, _vp]),
    "cfd_zpart_apply_xyz": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cfd_create_npts": (_i, [_pp, _i, _i, _i, _i, _d, _i, _i]),
    "cfd_zpart_apply_npts": (_i, [_vp, _vp, _vp, _vp]),
    "cfd_zpart_destroy": (None, [_vp]),
    "cfd_set_wait_timeout_ms": (_i, [_l]),
    "cfd_async_status": (_i, []),
    "cfd_plane_elems": (_l, [_vp]),
    "cfd_tables_size": (_i, []),
    "cfd_plan_tables": (_i, [_vp, _dp]),
    "cfd_debug_tables": (_i, [_i, _dp, _d, _dp]),
    "cfd_debug_xy_order": (_l, [_i, _i, _i, _d, _i, ctypes.POINTER(_i), _l]),
    "cfd_debug_xy_shape": (_i, [_i, _i, _i, _i, ctypes.POINTER(_i), _dp, ctypes.POINTER(_i)]),
    "cfd_debug_stream": (_i, [_vp, _vp, _vp, _l, _vp]),
    "cfd_debug_halo_weights": (_i, [_i, _d, _dp, _dp]),
    "cfd_debug_lookahead": (_i, [_i, _dp]),
    "cfd_debug_scheme": (_i, [_i, _i, _d, _dp]),
    "cfd_debug_secondary": (_i, [_i, _i, _i, _dp, _dp, _dp, _dp, _dp]),
    "cfd_plan_secondary": (_i, [_vp, _dp, _dp, _dp, _dp, _dp]),
    "nt_create": (_i, [_pp, _i, _i, _i, _i, _dp]),
    "nt_solve": (_i, [_vp, _vp, _vp]),
    "nt_destroy": (None, [_vp]),
    "nt_is_exact_two_pass": (_i, [_vp]),
    "nt_lookahead": (_i, [_vp]),
    "cfd_pthomas": (_i, [_dp, _dp, _dp, _vp, _i, _l, _vp]),
    "cfd_set_launch": (_i, [_i, _i, _i]),
    "cfd_set_segments": (_i, [_i]),
    "cfd_launch_count": (_l, []),
}


class CfdError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libcfd_b200 error {code}: {text}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m compact_finite_differences_b200.build` "
                "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the library is stale
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int) -> None:
    if code != CFD_OK:
        raise CfdError(code, lib().cfd_last_error().decode())
