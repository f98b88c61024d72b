"""ctypes binding of the fiber-emulated kernel (tests/emu/emu_driver.cpp).  TEST ONLY."""
import ctypes as C
import os
import subprocess

from path_optimizer_2_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libpqp_emu.so")
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, "emu_driver.cpp"), os.path.join(_HERE, "warp_emu.h"),
            os.path.join(_ROOT, "path_optimizer_2_b200", "csrc", "pqp_kernel.cuh"),
            os.path.join(_ROOT, "path_optimizer_2_b200", "csrc", "pqp_host_common.h")]
    stale = not os.path.exists(_LIB) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    if force or stale:
        subprocess.check_call(
            ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I", _HERE,
             "-I", os.path.join(_ROOT, "include"), "-o", _LIB, srcs[0]])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.POINTER(abi.PqpParams), C.c_int, C.c_int]
        L.emu_destroy.argtypes = [C.c_void_p]
        for f in (L.emu_solve, L.emu_resolve):
            f.argtypes = [C.c_void_p, C.POINTER(abi.PqpBatchIn), C.POINTER(abi.PqpBatchOut)]
        L.emu_chunk.argtypes = [C.c_void_p]
        L.emu_default_params.argtypes = [C.POINTER(abi.PqpParams)]
        _lib = L
    return _lib


class EmuSolver:
    def __init__(self, params, n_max, batch_max):
        self.L = lib()
        self.h = self.L.emu_create(C.byref(params), n_max, batch_max)
        if not self.h:
            raise RuntimeError("emu_create failed")
        self.n_max = n_max

    def __del__(self):
        if getattr(self, "h", None):
            self.L.emu_destroy(self.h)
            self.h = None

    def _run(self, fn, hb, full):
        res = abi.HostResult(hb.batch, hb.n_max, full=full)
        bi, bo = hb.as_struct(), res.as_struct()
        rc = fn(self.h, C.byref(bi), C.byref(bo))
        if rc:
            raise RuntimeError("emu rc=%d" % rc)
        return res

    def solve(self, hb, full=True):
        return self._run(self.L.emu_solve, hb, full)

    def resolve(self, hb, full=True):
        return self._run(self.L.emu_resolve, hb, full)
