"""ctypes binding of the clearance-bounds entry points (include/pqp_bounds.h).

`PathBounds` is the batched stand-in for `ReferencePathImpl::updateBoundsImproved`
(reference_path_impl.cpp:177-230): one handle owns the device copy of the map's distance layer;
`compute()` takes the splines and reference states of a batch of paths and returns the
front / rear / centre clearance bounds per state plus the index the reference would cut each
path at. There is no CPU path: without the CUDA library or a B200, construction raises.
"""
import ctypes as C

import numpy as np

from . import abi
from .solver import PqpError, load_library

SPLINE_ROWS, STATE_ROWS, BOUND_ROWS = 9, 4, 6

EXPORTED_SYMBOLS = [
    "pqp_bounds_default_params", "pqp_bounds_create", "pqp_bounds_destroy", "pqp_bounds_compute",
    "pqp_bounds_compute_device", "pqp_bounds_build_states", "pqp_bounds_build_states_device",
    "pqp_bounds_last_kernel_ms", "pqp_bounds_last_error",
]


class BoundsMap(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("resolution", C.c_double),
                ("center_x", C.c_double), ("center_y", C.c_double), ("distance", C.c_void_p)]


class BoundsParams(C.Structure):
    _fields_ = [("front_length", C.c_double), ("rear_length", C.c_double), ("car_width", C.c_double),
                ("safety_margin", C.c_double), ("epsilon", C.c_double)]


class BoundsIn(C.Structure):
    _fields_ = [("batch", C.c_int32), ("n_max", C.c_int32), ("k_max", C.c_int32), ("states", C.c_void_p),
                ("n", C.c_void_p), ("spline", C.c_void_p), ("k", C.c_void_p)]


class BoundsOut(C.Structure):
    _fields_ = [("bounds", C.c_void_p), ("n_valid", C.c_void_p), ("knots", C.c_void_p)]


class StatesIn(C.Structure):
    _fields_ = [("batch", C.c_int32), ("n_max", C.c_int32), ("k_max", C.c_int32), ("spline", C.c_void_p),
                ("k", C.c_void_p), ("max_s", C.c_void_p), ("delta_s_smaller", C.c_double),
                ("delta_s_larger", C.c_double), ("dynamic_segmentation", C.c_int32)]


class StatesOut(C.Structure):
    _fields_ = [("states", C.c_void_p), ("curvature", C.c_void_p), ("n", C.c_void_p), ("total", C.c_void_p),
                ("knots", C.c_void_p)]


def default_params(**overrides):
    p = BoundsParams(3.9, -1.0, 2.0, 0.3, 1e-6)
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def _declare(L):
    if getattr(L, "_pqp_bounds_declared", False):
        return L
    vp = C.c_void_p
    L.pqp_bounds_default_params.argtypes = [C.POINTER(BoundsParams)]
    L.pqp_bounds_default_params.restype = None
    L.pqp_bounds_create.argtypes = [C.POINTER(BoundsMap), C.POINTER(BoundsParams), C.c_int32, C.POINTER(vp)]
    L.pqp_bounds_destroy.argtypes = [vp]
    L.pqp_bounds_destroy.restype = None
    L.pqp_bounds_compute.argtypes = [vp, C.POINTER(BoundsIn), C.POINTER(BoundsOut)]
    L.pqp_bounds_compute_device.argtypes = [vp, C.POINTER(BoundsIn), C.POINTER(BoundsOut), vp]
    L.pqp_bounds_build_states.argtypes = [vp, C.POINTER(StatesIn), C.POINTER(StatesOut)]
    L.pqp_bounds_build_states_device.argtypes = [vp, C.POINTER(StatesIn), C.POINTER(StatesOut), vp]
    L.pqp_bounds_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.pqp_bounds_last_error.argtypes = [vp]
    L.pqp_bounds_last_error.restype = C.c_char_p
    L._pqp_bounds_declared = True
    return L


class PathBounds:
    """One obstacle map on one GPU; `dist` is the float32 distance layer [rows, cols]."""

    def __init__(self, dist, resolution, *, center=(0.0, 0.0), params=None, device=0):
        self.L = _declare(load_library())
        self.dist = np.ascontiguousarray(dist, dtype=np.float32)
        self.params = params if params is not None else default_params()
        m = BoundsMap(self.dist.shape[0], self.dist.shape[1], float(resolution), float(center[0]), float(center[1]),
                      self.dist.ctypes.data)
        h = C.c_void_p()
        rc = self.L.pqp_bounds_create(C.byref(m), C.byref(self.params), int(device), C.byref(h))
        if rc:
            raise PqpError(rc, (self.L.pqp_bounds_last_error(None) or b"").decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.pqp_bounds_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc):
        if rc:
            raise PqpError(rc, (self.L.pqp_bounds_last_error(self.h) or b"").decode())

    def compute(self, states, n, spline, k, *, knots=None):
        """Host buffers: states[b][4][n_max], n[b], spline[b][9][k_max], k[b] ->
        (bounds[b][6][n_max], n_valid[b]); `knots` (solver block [b][9][n_max]) is filled in place."""
        states = np.ascontiguousarray(states, dtype=np.float64)
        spline = np.ascontiguousarray(spline, dtype=np.float64)
        n = np.ascontiguousarray(n, dtype=np.int32)
        k = np.ascontiguousarray(k, dtype=np.int32)
        B, _, n_max = states.shape
        bounds = np.zeros((B, BOUND_ROWS, n_max))
        n_valid = np.zeros(B, dtype=np.int32)
        if knots is not None:
            assert knots.flags.c_contiguous and knots.dtype == np.float64 and knots.shape == (B, abi.NFIELDS, n_max)
        bi = BoundsIn(B, n_max, spline.shape[2], states.ctypes.data, n.ctypes.data, spline.ctypes.data, k.ctypes.data)
        bo = BoundsOut(bounds.ctypes.data, n_valid.ctypes.data, knots.ctypes.data if knots is not None else None)
        self._check(self.L.pqp_bounds_compute(self.h, C.byref(bi), C.byref(bo)))
        return bounds, n_valid

    def compute_device(self, bin_struct: BoundsIn, bout_struct: BoundsOut, stream=0):
        """Raw device pointers (e.g. torch tensors' data_ptr()), asynchronous on `stream`."""
        self._check(self.L.pqp_bounds_compute_device(self.h, C.byref(bin_struct), C.byref(bout_struct),
                                                     C.c_void_p(stream)))

    def build_states(self, spline, k, max_s, n_max, *, ds_small=0.15, ds_large=0.3, dynamic=True, knots=None):
        """buildReferenceFromSpline for a batch (host buffers): -> (states[b][4][n_max],
        curvature[b][n_max], n[b], total[b]); `knots` (solver block) is filled in place."""
        spline = np.ascontiguousarray(spline, dtype=np.float64)
        k = np.ascontiguousarray(k, dtype=np.int32)
        max_s = np.ascontiguousarray(max_s, dtype=np.float64)
        B = spline.shape[0]
        states, curv = np.zeros((B, STATE_ROWS, n_max)), np.zeros((B, n_max))
        n, total = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        if knots is not None:
            assert knots.flags.c_contiguous and knots.dtype == np.float64 and knots.shape == (B, abi.NFIELDS, n_max)
        si = StatesIn(B, n_max, spline.shape[2], spline.ctypes.data, k.ctypes.data, max_s.ctypes.data, ds_small, ds_large,
                      1 if dynamic else 0)
        so = StatesOut(states.ctypes.data, curv.ctypes.data, n.ctypes.data, total.ctypes.data,
                       knots.ctypes.data if knots is not None else None)
        self._check(self.L.pqp_bounds_build_states(self.h, C.byref(si), C.byref(so)))
        return states, curv, n, total

    def build_states_device(self, sin_struct: StatesIn, sout_struct: StatesOut, stream=0):
        self._check(self.L.pqp_bounds_build_states_device(self.h, C.byref(sin_struct), C.byref(sout_struct),
                                                          C.c_void_p(stream)))

    @property
    def last_kernel_ms(self):
        ms = C.c_float(0)
        self._check(self.L.pqp_bounds_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value
