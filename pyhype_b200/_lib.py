"""ctypes binding of the C ABI declared in ``include/pyh_b200.h``.

There is no CPU fallback: if the shared library is missing or fails to load, or no CUDA device
is visible, importing succeeds but the first engine call raises ``RuntimeError`` loudly.
"""
from __future__ import annotations

import ctypes as C
import os

PYH_ABI_VERSION = 1
PYH_MAX_STAGES = 6

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYH_LIB_PATH") or os.path.join(_HERE, "lib", "libpyh_b200.so")

c_double_p = C.POINTER(C.c_double)


class PyhConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("device", C.c_int32),
        ("nx", C.c_int32),
        ("ny", C.c_int32),
        ("flux", C.c_int32),
        ("limiter", C.c_int32),
        ("recon", C.c_int32),
        ("num_quadrature_points", C.c_int32),
        ("num_stages", C.c_int32),
        ("reserved", C.c_int32),
        ("tableau", C.c_double * (PYH_MAX_STAGES * PYH_MAX_STAGES)),
        ("gamma", C.c_double),
        ("cfl", C.c_double),
    ]


class PyhBlockDesc(C.Structure):
    _fields_ = [
        ("gid", C.c_int32),
        ("is_cartesian", C.c_int32),
        ("neighbor", C.c_int32 * 4),
        ("neighbor_is_local", C.c_int32 * 4),
        ("bc", C.c_int32 * 4),
        ("nodes_x", c_double_p),
        ("nodes_y", c_double_p),
        ("area", c_double_p),
        ("cos_v", c_double_p),
        ("sin_v", c_double_p),
        ("cos_h", c_double_p),
        ("sin_h", c_double_p),
        ("dirichlet_prim", c_double_p * 4),
    ]


# every symbol include/pyh_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SIGNATURES = {
    "pyh_last_error": (C.c_char_p, []),
    "pyh_abi_version": (C.c_int, []),
    "pyh_create": (C.c_int, [C.POINTER(PyhConfig), C.POINTER(_vp)]),
    "pyh_add_block": (C.c_int, [_vp, C.POINTER(PyhBlockDesc)]),
    "pyh_finalize": (C.c_int, [_vp]),
    "pyh_destroy": (C.c_int, [_vp]),
    "pyh_upload_state": (C.c_int, [_vp, C.c_int, c_double_p]),
    "pyh_download_state": (C.c_int, [_vp, C.c_int, c_double_p]),
    "pyh_fill_uniform": (C.c_int, [_vp, C.c_int, c_double_p]),
    "pyh_fill_box": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, c_double_p, c_double_p]),
    "pyh_upload_state_async": (C.c_int, [_vp, C.c_int, c_double_p]),
    "pyh_commit_uploads": (C.c_int, [_vp]),
    "pyh_download_state_async": (C.c_int, [_vp, C.c_int, c_double_p]),
    "pyh_transfers_sync": (C.c_int, [_vp]),
    "pyh_downloads_sync": (C.c_int, [_vp]),
    "pyh_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "pyh_host_free": (C.c_int, [_vp]),
    "pyh_download_ghost": (C.c_int, [_vp, C.c_int, C.c_int, c_double_p]),
    "pyh_apply_bc": (C.c_int, [_vp]),
    "pyh_halo_count": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "pyh_halo_slot": (
        C.c_int,
        [_vp, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
         C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    ),
    "pyh_pack_halo": (C.c_int, [_vp, _vp]),
    "pyh_unpack_halo": (C.c_int, [_vp, _vp]),
    "pyh_local_dt": (C.c_int, [_vp, _vp]),
    "pyh_get_dt": (C.c_int, [_vp, C.c_double, C.c_double, c_double_p]),
    "pyh_step_begin": (C.c_int, [_vp, C.c_double]),
    "pyh_step_begin_dev": (C.c_int, [_vp, _vp]),
    "pyh_stage": (C.c_int, [_vp, C.c_int]),
    "pyh_step": (C.c_int, [_vp, C.c_double]),
    "pyh_run": (
        C.c_int,
        [_vp, c_double_p, C.c_double, C.c_int64, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
         c_double_p, C.c_int64],
    ),
    "pyh_realizable": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "pyh_comm_unique_id": (C.c_int, [_vp]),
    "pyh_comm_init": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, C.POINTER(C.c_int32), C.c_int32]),
    "pyh_comm_info": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "pyh_residual": (C.c_int, [_vp, C.c_int, c_double_p]),
    "pyh_debug_fetch": (C.c_int, [_vp, C.c_int, C.c_int, c_double_p]),
    "pyh_march_shape": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pyh_stage_path": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "pyh_launch_count": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "pyh_stream": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "pyh_sync": (C.c_int, [_vp]),
}

_lib = None


def load():
    """Load ``libpyh_b200.so`` (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"pyhype_b200: CUDA library not built ({LIB_PATH} missing). Run `python -c 'import "
            "__graft_entry__ as g; g.build()'` -- there is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.pyh_abi_version() != PYH_ABI_VERSION:
        raise RuntimeError("pyhype_b200: ABI version mismatch between _lib.py and libpyh_b200.so")
    _lib = lib
    return lib


class PyhError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        msg = load().pyh_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise PyhError(f"pyh error {rc}: {msg}")
