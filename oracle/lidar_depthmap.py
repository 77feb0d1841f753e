"""ctypes front of oracle/c/lidar_depthmap.c (TEST INFRASTRUCTURE ONLY; see its header: generate_depth
is PARITY UNPINNED -- the reference needs Eigen / OpenCV / pybind11 and cannot run here -- minpool is
pinned to the reference's Python through tests/golden/lidar.npz)."""
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


def minpool(x, scale, default=0.0):
    x = np.ascontiguousarray(x, dtype=np.float32)
    H, W = x.shape
    out = np.zeros((H // scale, W // scale), dtype=np.float32)
    if default:
        _load().oracle_minpool(_fp(x), W, H, scale, default, _fp(out))
    else:   # plain minimum (utils/img_utils.py:93-94)
        out = x[:H // scale * scale, :W // scale * scale].reshape(H // scale, scale, W // scale, scale).min((1, 3))
    return out
