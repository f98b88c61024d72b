"""ctypes binding of tests/emu/smoother_driver.cpp (the smoother-QP routine of the CUDA library on the CPU). TEST ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "libpqs_emu.so")
_lib = None
# rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance, max_iter, check, scaling,
# adaptive_rho, adaptive_rho_interval  (OSQP 0.6.x defaults; interval fixed at 25 like the oracle)
DEFAULT_SETTINGS = (0.1, 1e-6, 1.6, 1e-3, 1e-3, 1e-4, 1e-4, 5.0, 4000, 25, 10, 1, 25)


def _load():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "smoother_driver.cpp"),
                os.path.join(_ROOT, "path_optimizer_2_b200", "csrc", "pqp_smoother_core.cuh")]
        if not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", _LIB, srcs[0]])
        _lib = C.CDLL(_LIB)
    return _lib


def _arr(v):
    return np.ascontiguousarray(v, dtype=np.float64)


def tension(x, y, a, k, s, weights=(0.005, 1.0, 10.0), settings=DEFAULT_SETTINGS):
    L = _load()
    x, y, a, k, s = (_arr(v) for v in (x, y, a, k, s))
    p = len(x)
    rx, ry, rs, xf, info = np.zeros(p), np.zeros(p), np.zeros(p), np.zeros(4 * p - 1), np.zeros(6)
    w, st = _arr(weights), _arr(settings)
    L.smoother_emu_tension.argtypes = [C.c_int] + [C.c_void_p] * 12
    L.smoother_emu_tension(p, *(v.ctypes.data for v in (x, y, a, k, s, w, st, rx, ry, rs, xf, info)))
    return dict(status=int(info[0]), iters=int(info[1]), rho_updates=int(info[2]), pri_res=info[3], dua_res=info[4], obj=info[5],
                x=rx, y=ry, s=rs, x_full=xf)


def post(layer_s, lower, upper, vehicle_l, weights=(1.0, 100.0, 1000.0), settings=DEFAULT_SETTINGS):
    L = _load()
    layer_s, lower, upper = (_arr(v) for v in (layer_s, lower, upper))
    p = len(layer_s)
    off, xf, info = np.zeros(p), np.zeros(3 * p), np.zeros(6)
    w, st = _arr(weights), _arr(settings)
    L.smoother_emu_post.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 5
    L.smoother_emu_post(p, layer_s.ctypes.data, lower.ctypes.data, upper.ctypes.data, float(vehicle_l),
                        *(v.ctypes.data for v in (w, st, off, xf, info)))
    return dict(status=int(info[0]), iters=int(info[1]), rho_updates=int(info[2]), pri_res=info[3], dua_res=info[4], obj=info[5],
                offsets=off, x_full=xf)
