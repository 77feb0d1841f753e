"""ctypes binding of libdpv_sm100a.so (include/dpv_b200.h).

The library is the product; there is no Python or CPU fallback.  If the shared
object is missing it is built in-tree with nvcc (build.py); if that is not
possible the import of any op raises.
"""
import ctypes
import os
import threading

from . import build as _build

_c_fp = ctypes.c_void_p     # device pointers travel as integers
_c_i = ctypes.c_int
_c_i64 = ctypes.c_int64
_c_f = ctypes.c_float

# name -> (restype, argtypes); mirrors include/dpv_b200.h one to one.
PROTOTYPES = {
    "dpv_abi_version": (_c_i, []),
    "dpv_error_string": (ctypes.c_char_p, [_c_i]),
    "dpv_launch_count": (ctypes.c_longlong, []),
    "dpv_sweep_cost_volume": (_c_i, [_c_fp] * 8 + [_c_i] * 6 + [_c_i64] * 6 + [_c_f, _c_i, _c_i, _c_fp]),
    "dpv_sweep_workspace_floats": (_c_i64, [_c_i] * 4),
    "dpv_sweep_cost_volume_ws": (_c_i, [_c_fp] * 8 + [_c_i] * 6 + [_c_i64] * 6 + [_c_f, _c_i, _c_i, _c_fp, _c_fp]),
    "dpv_warp_planes": (_c_i, [_c_fp] * 5 + [_c_i] * 4 + [_c_i64, _c_f, _c_f, _c_fp]),
    "dpv_warp_feature": (_c_i, [_c_fp] * 6 + [_c_i] * 5 + [_c_i64] * 3 + [_c_fp]),
    "dpv_head": (_c_i, [_c_fp] * 9 + [_c_i] * 5 + [_c_fp]),
    "dpv_lidar_prior": (_c_i, [_c_fp] * 4 + [_c_i] * 4 + [_c_f, _c_fp]),
    "dpv_bayes_fuse": (_c_i, [_c_fp] * 7 + [_c_i] * 4 + [_c_f, _c_fp]),
    "dpv_ufield_workspace_floats": (_c_i64, [_c_i] * 4),
    "dpv_ufield": (_c_i, [_c_fp] * 12 + [_c_i] * 4 + [_c_i64, _c_i] + [_c_f] * 6 + [_c_fp]),
    "dpv_head_ufield_workspace_floats": (_c_i64, [_c_i] * 4),
    "dpv_uf_fused_tables": (_c_i, [_c_fp] * 4 + [_c_i] * 2 + [_c_fp] * 2),
    "dpv_head_ufield": (_c_i, [_c_fp] * 13 + [_c_i] * 4 + [_c_i64, _c_i] + [_c_f] * 5 + [_c_fp]),
    "dpv_correlation": (_c_i, [_c_fp] * 3 + [_c_i] * 5 + [_c_fp]),
    "dpv_shard_stats": (_c_i, [_c_fp] * 3 + [_c_i] * 5 + [_c_fp]),
    "dpv_shard_merge_slice": (_c_i, [_c_fp] * 2 + [_c_i] * 2 + [_c_fp]),
    "dpv_shard_finish": (_c_i, [_c_fp] * 6 + [_c_i] * 4 + [_c_fp]),
    "dpv_shard_merge_finish": (_c_i, [_c_fp] * 6 + [_c_i] * 4 + [_c_fp]),
    "dpv_conv3x3_packed_floats": (_c_i64, [_c_i] * 3),
    "dpv_conv3x3_pack": (_c_i, [_c_fp] * 3 + [_c_i] * 4 + [_c_fp]),
    "dpv_conv3x3_pack_weights": (_c_i, [_c_fp] * 3 + [_c_i] * 2 + [_c_fp]),
    "dpv_conv3x3_d64": (_c_i, [_c_fp] * 8 + [_c_i] * 4 + [_c_f, _c_fp]),
    "dpv_conv3d_packed_floats": (_c_i64, [_c_i] * 4),
    "dpv_conv3d_pack": (_c_i, [_c_fp] * 3 + [_c_i] * 5 + [_c_fp]),
    "dpv_conv3d_pack_weights": (_c_i, [_c_fp] * 4 + [_c_i] * 2 + [_c_fp]),
    "dpv_conv3d_c32": (_c_i, [_c_fp] * 12 + [_c_i] * 6 + [_c_fp]),
    "dpv_conv3d_pack_zfold": (_c_i, [_c_fp] * 3 + [_c_i] * 5 + [_c_fp]),
    "dpv_conv3d_pack_weights_zfold": (_c_i, [_c_fp] * 4 + [_c_i] * 2 + [_c_fp]),
    "dpv_conv3d_c32_to1": (_c_i, [_c_fp] * 4 + [_c_i] * 5 + [_c_fp]),
    "dpv_conv3d_bn_apply": (_c_i, [_c_fp] * 4 + [_c_f] + [_c_fp] * 4 + [_c_i] * 5 + [_c_fp]),
    "dpv_depth_errors_workspace_doubles": (_c_i64, [_c_i] * 3),
    "dpv_depth_errors": (_c_i, [_c_fp] * 3 + [_c_f, _c_i] + [_c_fp] * 3 + [_c_i] * 3 + [_c_fp]),
    "dpv_unc_rmse": (_c_i, [_c_fp] * 4 + [_c_i] * 3 + [_c_fp]),
    "dpv_lidar_depthmap": (_c_i, [_c_fp, _c_i, _c_fp, _c_fp] + [_c_i] * 3 + [_c_f, _c_i, _c_f] + [_c_fp] * 5),
    "dpv_minpool": (_c_i, [_c_fp, _c_fp] + [_c_i] * 4 + [_c_f, _c_fp]),
    "dpv_pipeline_create": (_c_i, [ctypes.POINTER(ctypes.c_void_p)] + [_c_i] * 9),
    "dpv_pipeline_destroy": (_c_i, [ctypes.c_void_p]),
    "dpv_pipeline_run": (_c_i, [ctypes.c_void_p] + [_c_fp] * 11 + [_c_f] + [_c_fp] * 7),
    "dpv_pipeline_submit": (_c_i, [ctypes.c_void_p] + [_c_fp] * 11 + [_c_f] + [_c_fp] * 7),
    "dpv_pipeline_wait": (_c_i, [ctypes.c_void_p]),
    "dpv_pipeline_last_bytes": (_c_i, [ctypes.c_void_p, ctypes.POINTER(_c_i64),
                                       ctypes.POINTER(_c_i64)]),
}

_lock = threading.Lock()
_lib = None


class DpvError(RuntimeError):
    pass


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first if needed) the shared library; raise if impossible."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        if not os.path.exists(path):
            path = _build.build()
        lib = ctypes.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError = ABI mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(code):
    if code != 0:
        msg = load().dpv_error_string(int(code))
        raise DpvError("libdpv_sm100a: %s (code %d)" % (msg.decode() if msg else "?", code))


def launch_count():
    return int(load().dpv_launch_count())
