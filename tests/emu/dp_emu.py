"""ctypes binding of tests/emu/dp_driver.cpp (the DP routine of the CUDA library on the CPU). TEST ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "libpqd_emu.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, "dp_driver.cpp"), os.path.join(_ROOT, "path_optimizer_2_b200", "csrc", "pqp_dp_core.cuh"),
            os.path.join(_ROOT, "path_optimizer_2_b200", "csrc", "pqp_bounds_core.cuh")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", _LIB, srcs[0]])
    return _LIB


def search(dist, res, spline, k, length, start, layers_max=160, params=(10.0, 0.6, 1.5, 2.0)):
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
    from path_optimizer_2_b200 import dp
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    spline = np.ascontiguousarray(spline, dtype=np.float64)
    k = np.ascontiguousarray(k, dtype=np.int32)
    length = np.ascontiguousarray(length, dtype=np.float64)
    start = np.ascontiguousarray(start, dtype=np.float64)
    B, k_max = spline.shape[0], spline.shape[2]
    J = 0
    cl = -params[0]
    while cl <= params[0]:
        J += 1
        cl += params[1]
    r = dp.DpResult(B, layers_max, J)
    vp = C.c_void_p
    d = C.c_double
    _lib.dp_emu_search.argtypes = [vp, C.c_int, C.c_int, d, d, d, d, d, C.c_int, C.c_int, C.c_int] + [vp] * 16
    got = _lib.dp_emu_search(dist.ctypes.data, dist.shape[0], dist.shape[1], res, *params, B, k_max, layers_max,
                             spline.ctypes.data, k.ctypes.data, length.ctypes.data, start.ctypes.data, r.status.ctypes.data,
                             r.n_layers.ctypes.data, r.n_out.ctypes.data, r.layer_s.ctypes.data, r.lower.ctypes.data,
                             r.upper.ctypes.data, r.chosen.ctypes.data, r.vehicle_l.ctypes.data, r.target_s.ctypes.data,
                             r.cost.ctypes.data, r.parent.ctypes.data, r.feasible.ctypes.data)
    assert got == J
    return r
