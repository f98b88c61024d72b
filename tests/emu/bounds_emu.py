"""ctypes binding of tests/emu/bounds_driver.cpp (the bounds kernel's source compiled for the host). TEST ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from path_optimizer_2_b200 import bounds as pb

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libpqb_emu.so")
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "bounds_driver.cpp"),
                os.path.join(_ROOT, "path_optimizer_2_b200", "csrc", "pqp_bounds_core.cuh"),
                os.path.join(_ROOT, "include", "pqp_bounds.h")]
        if not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
                                   "-I", os.path.join(_ROOT, "include"), "-o", _LIB, srcs[0]])
        _lib = C.CDLL(_LIB)
        _lib.pqb_emu_compute.argtypes = [C.POINTER(pb.BoundsMap), C.POINTER(pb.BoundsParams), C.POINTER(pb.BoundsIn),
                                         C.POINTER(pb.BoundsOut)]
        _lib.pqb_emu_build_states.argtypes = [C.POINTER(pb.StatesIn), C.POINTER(pb.StatesOut)]
    return _lib


def compute(dist, res, states, n, spline, k, *, params=None, center=(0.0, 0.0), knots=None):
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    states = np.ascontiguousarray(states, dtype=np.float64)
    spline = np.ascontiguousarray(spline, dtype=np.float64)
    n = np.ascontiguousarray(n, dtype=np.int32)
    k = np.ascontiguousarray(k, dtype=np.int32)
    B, _, n_max = states.shape
    out = np.zeros((B, 6, n_max))
    n_valid = np.zeros(B, dtype=np.int32)
    m = pb.BoundsMap(dist.shape[0], dist.shape[1], res, center[0], center[1], dist.ctypes.data)
    p = params if params is not None else pb.default_params()
    bi = pb.BoundsIn(B, n_max, spline.shape[2], states.ctypes.data, n.ctypes.data, spline.ctypes.data, k.ctypes.data)
    bo = pb.BoundsOut(out.ctypes.data, n_valid.ctypes.data, knots.ctypes.data if knots is not None else None)
    assert lib().pqb_emu_compute(C.byref(m), C.byref(p), C.byref(bi), C.byref(bo)) == 0
    return out, n_valid


def build_states(spline, k, max_s, n_max, *, ds_small=0.15, ds_large=0.3, dynamic=True):
    spline = np.ascontiguousarray(spline, dtype=np.float64)
    k = np.ascontiguousarray(k, dtype=np.int32)
    max_s = np.ascontiguousarray(max_s, dtype=np.float64)
    B = spline.shape[0]
    states, curv = np.zeros((B, 4, n_max)), np.zeros((B, n_max))
    n, total = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    si = pb.StatesIn(B, n_max, spline.shape[2], spline.ctypes.data, k.ctypes.data, max_s.ctypes.data, ds_small, ds_large,
                     1 if dynamic else 0)
    so = pb.StatesOut(states.ctypes.data, curv.ctypes.data, n.ctypes.data, total.ctypes.data, None)
    assert lib().pqb_emu_build_states(C.byref(si), C.byref(so)) == 0
    return states, curv, n, total
