"""ctypes binding of include/pqp_multi.h: one process, the GPUs of one box (contiguous shards, one
NCCL all-gather of 16 B per instance). `MultiGpuSolver.solve/resolve` mirror `PathQpSolver`'s."""
import ctypes as C

import numpy as np

from . import abi, solver

EXPORTED_SYMBOLS = [
    "pqp_multi_create", "pqp_multi_destroy", "pqp_multi_solve", "pqp_multi_resolve", "pqp_multi_gathered",
    "pqp_multi_shard", "pqp_multi_handle", "pqp_multi_last_gather_ms", "pqp_multi_last_error",
    "pqp_pack_results_device", "pqp_resident_results",
]

RESULT_REC = np.dtype([("cost", np.float64), ("status", np.int32), ("iters", np.int32)])  # pqp_result_rec


def _lib():
    L = solver.load_library()
    if not getattr(L, "_multi_declared", False):
        vp = C.c_void_p
        pin, pout = C.POINTER(abi.PqpBatchIn), C.POINTER(abi.PqpBatchOut)
        L.pqp_multi_create.argtypes = [C.POINTER(abi.PqpParams), C.c_int32, C.c_int32, C.c_int32, vp, C.POINTER(vp)]
        L.pqp_multi_destroy.argtypes = [vp]
        L.pqp_multi_solve.argtypes = [vp, pin, pout]
        L.pqp_multi_resolve.argtypes = [vp, pin, pout]
        L.pqp_multi_gathered.argtypes = [vp, C.c_int32, C.POINTER(vp), C.POINTER(C.c_int32)]
        L.pqp_multi_shard.argtypes = [vp, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.pqp_multi_handle.argtypes = [vp, C.c_int32, C.POINTER(vp)]
        L.pqp_multi_last_gather_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.pqp_multi_last_error.argtypes = [vp]
        L.pqp_multi_last_error.restype = C.c_char_p
        L.pqp_pack_results_device.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
        L.pqp_resident_results.argtypes = [vp] + [C.POINTER(vp)] * 4
        L._multi_declared = True
    return L


def pack_results_device(sv: "solver.PathQpSolver", batch, cost_ptr, status_ptr, iters_ptr, packed_ptr, stream=0):
    """{cost, status, iters} -> 16-byte records, one launch (pqp_pack_results_device)."""
    sv._check(_lib().pqp_pack_results_device(sv.h, batch, cost_ptr, status_ptr, iters_ptr, packed_ptr, C.c_void_p(stream)))


class MultiGpuSolver:
    def __init__(self, params=None, *, n_max, batch_max, n_devices, devices=None):
        self.L = _lib()
        self.params = params if params is not None else abi.default_params()
        self.n_max, self.batch_max, self.n_devices = int(n_max), int(batch_max), int(n_devices)
        dv = None
        if devices is not None:
            dv = (C.c_int32 * self.n_devices)(*devices)
        m = C.c_void_p()
        rc = self.L.pqp_multi_create(C.byref(self.params), self.n_max, self.batch_max, self.n_devices, dv, C.byref(m))
        if rc:
            raise solver.PqpError(rc, (self.L.pqp_multi_last_error(None) or b"").decode())
        self.m = m

    def close(self):
        if getattr(self, "m", None):
            self.L.pqp_multi_destroy(self.m)
            self.m = None

    __del__ = close

    def _check(self, rc):
        if rc:
            raise solver.PqpError(rc, (self.L.pqp_multi_last_error(self.m) or b"").decode())

    def solve(self, hb: abi.HostBatch, *, full=False, out: abi.HostResult = None) -> abi.HostResult:
        res = out if out is not None else abi.HostResult(hb.batch, hb.n_max, full=full)
        bi, bo = hb.as_struct(), res.as_struct()
        self._check(self.L.pqp_multi_solve(self.m, C.byref(bi), C.byref(bo)))
        return res

    def resolve(self, hb: abi.HostBatch = None, *, full=False, batch=None, out: abi.HostResult = None) -> abi.HostResult:
        if hb is None:
            res = out if out is not None else abi.HostResult(batch, self.n_max, full=full)
            bo = res.as_struct()
            self._check(self.L.pqp_multi_resolve(self.m, None, C.byref(bo)))
            return res
        res = out if out is not None else abi.HostResult(hb.batch, hb.n_max, full=full)
        bi, bo = hb.as_struct(), res.as_struct()
        self._check(self.L.pqp_multi_resolve(self.m, C.byref(bi), C.byref(bo)))
        return res

    def gathered_ptr(self, device_index):
        """(device pointer of the gathered table on that device, records per device)."""
        p, per = C.c_void_p(), C.c_int32()
        self._check(self.L.pqp_multi_gathered(self.m, device_index, C.byref(p), C.byref(per)))
        return p.value, per.value

    @property
    def last_gather_ms(self):
        ms = C.c_float()
        self._check(self.L.pqp_multi_last_gather_ms(self.m, C.byref(ms)))
        return ms.value
