"""ctypes front of oracle/c/lidar_depthmap.c (TEST INFRASTRUCTURE ONLY) and of the reference's own
generate_depth compiled into oracle/_ref/libutils_ref.so (oracle/ref_utils_lib/: the reference's
external/utils_lib/python/utils_lib.cpp from where it lies, against minimal Eigen / OpenCV / pybind11
stand-ins -- those libraries are not in this image).  The restatement is pinned to that build bit for bit
(tests/test_lidar.py); what the stand-in has to choose -- the association of the 4-term sums in Eigen's two
small matrix products -- is stated in oracle/ref_utils_lib/stub/Eigen/Dense.  minpool is pinned to the
reference's Python through tests/golden/lidar.npz."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_c.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "c", "lidar_depthmap.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.run(["make", "-C", os.path.join(_HERE, "c")], check=True, stdout=subprocess.DEVNULL)
        lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        lib.oracle_generate_depth.restype = None
        lib.oracle_generate_depth.argtypes = [fp, ctypes.c_int, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_float, fp]
        lib.oracle_minpool.restype = None
        lib.oracle_minpool.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, fp]
        _lib = lib
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def generate_depth(velo, intr, m_velo2cam, width, height, filtering=2, filterdiff=1.0):
    velo = np.ascontiguousarray(velo, dtype=np.float32)
    intr = np.ascontiguousarray(intr, dtype=np.float32)
    m = np.ascontiguousarray(m_velo2cam, dtype=np.float32)
    out = np.zeros((height, width), dtype=np.float32)
    _load().oracle_generate_depth(_fp(velo), velo.shape[0], _fp(intr), _fp(m), width, height, filtering,
                                  filterdiff, _fp(out))
    return out


_REF_SO = os.path.join(_HERE, "_ref", "libutils_ref.so")
_ref = None


def reference_available():
    return os.path.exists(_REF_SO)


def reference_generate_depth(velo, intr, m_velo2cam, width, height, filtering=2, filterdiff=1.0):
    """The reference's own generate_depth (utils_lib.cpp:86-160, upsample = 0), compiled by oracle/ref_utils_lib."""
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(_REF_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        lib.ref_generate_depth.restype = ctypes.c_int
        lib.ref_generate_depth.argtypes = [fp, ctypes.c_int, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_float, fp]
        _ref = lib
    velo = np.ascontiguousarray(velo, dtype=np.float32)
    intr = np.ascontiguousarray(intr, dtype=np.float32)
    m = np.ascontiguousarray(m_velo2cam, dtype=np.float32)
    out = np.zeros((height, width), dtype=np.float32)
    rc = _ref.ref_generate_depth(_fp(velo), velo.shape[0], _fp(intr), _fp(m), width, height, filtering, filterdiff,
                                 _fp(out))
    if rc != 0:
        raise RuntimeError("ref_generate_depth failed")
    return out


def minpool(x, scale, default=0.0):
    x = np.ascontiguousarray(x, dtype=np.float32)
    H, W = x.shape
    out = np.zeros((H // scale, W // scale), dtype=np.float32)
    if default:
        _load().oracle_minpool(_fp(x), W, H, scale, default, _fp(out))
    else:   # plain minimum (utils/img_utils.py:93-94)
        out = x[:H // scale * scale, :W // scale * scale].reshape(H // scale, scale, W // scale, scale).min((1, 3))
    return out
