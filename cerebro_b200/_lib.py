"""ctypes binding of ``libcerebro_b200.so`` (the C ABI declared in include/cerebro_b200.h).

There is no CPU fallback: if the library is missing this module raises at import of the
symbol table, and every ``create`` call fails without a B200.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_native", "libcerebro_b200.so")


class CerebroB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("cerebro_b200 error %d: %s" % (code, msg))
        self.code = code


class NetvladWeights(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int),
        ("n_blocks", C.c_int),
        ("conv1_w", C.POINTER(C.c_float)),
        ("conv1_b", C.POINTER(C.c_float)),
        ("dw_w", C.POINTER(C.POINTER(C.c_float))),
        ("dw_b", C.POINTER(C.POINTER(C.c_float))),
        ("pw_w", C.POINTER(C.POINTER(C.c_float))),
        ("pw_b", C.POINTER(C.POINTER(C.c_float))),
        ("dw_stride", C.POINTER(C.c_int)),
        ("channels_out", C.POINTER(C.c_int)),
        ("vlad_k", C.c_int),
        ("vlad_d", C.c_int),
        ("vlad_w", C.POINTER(C.c_float)),
        ("vlad_b", C.POINTER(C.c_float)),
        ("vlad_c", C.POINTER(C.c_float)),
        ("vlad_ghost", C.c_int),
    ]


class IrBlock(C.Structure):  # cb_ir_block
    _fields_ = [
        ("c_in", C.c_int),
        ("c_exp", C.c_int),
        ("c_out", C.c_int),
        ("stride", C.c_int),
        ("residual", C.c_int),
        ("expand_w", C.POINTER(C.c_float)),
        ("expand_b", C.POINTER(C.c_float)),
        ("dw_w", C.POINTER(C.c_float)),
        ("dw_b", C.POINTER(C.c_float)),
        ("project_w", C.POINTER(C.c_float)),
        ("project_b", C.POINTER(C.c_float)),
    ]


class NetvladV2Weights(C.Structure):  # cb_netvlad_v2_weights
    _fields_ = [
        ("in_channels", C.c_int),
        ("conv1_w", C.POINTER(C.c_float)),
        ("conv1_b", C.POINTER(C.c_float)),
        ("n_blocks", C.c_int),
        ("blocks", C.POINTER(IrBlock)),
        ("vlad_k", C.c_int),
        ("vlad_d", C.c_int),
        ("vlad_w", C.POINTER(C.c_float)),
        ("vlad_b", C.POINTER(C.c_float)),
        ("vlad_c", C.POINTER(C.c_float)),
        ("vlad_ghost", C.c_int),
    ]


class RansacParams(C.Structure):
    _fields_ = [
        ("error_thresh", C.c_double),
        ("min_inlier_ratio", C.c_double),
        ("max_iterations", C.c_int),
        ("min_iterations", C.c_int),
        ("use_mle", C.c_int),
        ("failure_probability", C.c_double),
        ("adaptive", C.c_int),
        ("seed", C.c_uint64),
    ]


_vp = C.c_void_p
_i64 = C.c_int64
_SIGNATURES = {
    "cb_version": (C.c_int, []),
    "cb_last_error": (C.c_char_p, []),
    "cb_device_count": (C.c_int, []),
    "cb_index_create": (C.c_int, [C.POINTER(_vp), C.c_int, _i64, C.c_int, C.c_int, C.c_int]),
    "cb_index_destroy": (C.c_int, [_vp]),
    "cb_index_reset": (C.c_int, [_vp]),
    "cb_index_ntotal": (_i64, [_vp]),
    "cb_index_nlocal": (_i64, [_vp]),
    "cb_index_dim": (C.c_int, [_vp]),
    "cb_index_add": (C.c_int, [_vp, _i64, _vp]),
    "cb_index_add_f64": (C.c_int, [_vp, _i64, _vp]),
    "cb_index_add_device": (C.c_int, [_vp, _i64, _vp, _vp]),
    "cb_index_add_local_device": (C.c_int, [_vp, _i64, _vp, _vp]),
    "cb_index_set_timing": (C.c_int, [_vp, C.c_int]),
    "cb_index_get_sweep_timing": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "cb_index_search": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _i64, C.c_int, _vp, _vp, _vp]),
    "cb_index_search_device": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _i64, C.c_int, _vp, _vp, _vp]),
    "cb_topk_merge_device": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "cb_comm_get_unique_id": (C.c_int, [_vp]),
    "cb_comm_create": (C.c_int, [C.POINTER(_vp), _vp, C.c_int, C.c_int, C.c_int]),
    "cb_comm_destroy": (C.c_int, [_vp]),
    "cb_comm_rank": (C.c_int, [_vp]),
    "cb_comm_world": (C.c_int, [_vp]),
    "cb_comm_nccl_version": (C.c_int, []),
    "cb_index_attach_comm": (C.c_int, [_vp, _vp]),
    "cb_index_search_sharded_device": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _i64, C.c_int, _vp, _vp, _vp]),
    "cb_index_search_sharded": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _i64, C.c_int, _vp, _vp, _vp]),
    "cb_index_gathered_queries": (_vp, [_vp]),
    "cb_index_naive_candidate": (
        C.c_int,
        [_vp, _i64, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_int), C.POINTER(_i64), C.POINTER(C.c_double), C.POINTER(_i64)],
    ),
    "cb_index_get_rows": (C.c_int, [_vp, _i64, _i64, _vp]),
    "cb_index_device_rows": (_vp, [_vp]),
    "cb_frontend_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int]),
    "cb_frontend_destroy": (C.c_int, [_vp]),
    "cb_frontend_match_gms": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "cb_frontend_last_match_ms": (C.c_float, [_vp]),
    "cb_frontend_stereo_bm": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "cb_frontend_last_stereo_ms": (C.c_float, [_vp]),
    "cb_frontend_disparity_to_3d": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _vp]),
    "cb_frontend_collect": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cb_features_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cb_features_destroy": (C.c_int, [_vp]),
    "cb_features_set_remap": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "cb_features_remap": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, _vp]),
    "cb_features_orb": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cb_features_last_orb_ms": (C.c_float, [_vp]),
    "cb_features_debug_read": (_i64, [_vp, C.c_int, _vp, _i64]),
    "cb_descriptor_create": (C.c_int, [C.POINTER(_vp), C.POINTER(NetvladWeights), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cb_descriptor_create_v2": (C.c_int, [C.POINTER(_vp), C.POINTER(NetvladV2Weights), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cb_descriptor_destroy": (C.c_int, [_vp]),
    "cb_descriptor_dim": (C.c_int, [_vp]),
    "cb_descriptor_compute": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp]),
    "cb_descriptor_compute_f64": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp]),
    "cb_descriptor_compute_device": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "cb_descriptor_get_activation": (_i64, [_vp, C.c_int, _vp, _i64]),
    "cb_ransac_params_default": (None, [C.POINTER(RansacParams)]),
    "cb_pnp_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int]),
    "cb_pnp_destroy": (C.c_int, [_vp]),
    "cb_pnp_solve_batch": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.POINTER(RansacParams), _vp, _vp, _vp, _vp, _vp, _vp]),
    "cb_pnp_solve_batch_device": (
        C.c_int,
        [_vp, C.c_int, _vp, C.c_int, _vp, _vp, C.POINTER(RansacParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    ),
    "cb_pnp_icp_batch": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.POINTER(RansacParams), _vp, _vp, _vp, _vp, _vp, _vp]),
    "cb_pnp_icp_batch_device": (
        C.c_int,
        [_vp, C.c_int, _vp, _vp, _vp, C.POINTER(RansacParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    ),
    "cb_pnp_dls_minimal": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "cb_pnp_debug_read": (C.c_int64, [_vp, C.c_int, C.c_int, _vp, C.c_int64]),
}

_lib = None


def load():
    """Load the shared library (once) and attach signatures.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CerebroB200Error(
            -2,
            "native library %s not found: run `python -m cerebro_b200.build` (there is no CPU fallback)" % LIB_PATH,
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and this table disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def declared_symbols():
    return list(_SIGNATURES.keys())


def check(rc: int):
    if rc != 0:
        raise CerebroB200Error(rc, load().cb_last_error().decode("utf8", "replace"))


def ptr(a):
    """Pointer (as int) of a numpy array or a torch tensor, or None."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


def current_stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream
