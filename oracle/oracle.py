"""ctypes binding of the CPU oracle (oracle/pqp_oracle.c) + numpy checkers.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. Never imported by the product package.

PARITY UNPINNED (see oracle/pqp_oracle.h): there is no reference-produced number to pin
this oracle against; `kkt_report` below is the independent check that its answers are the
minimisers of the QP the reference assembles.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

from path_optimizer_2_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpqp_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "pqp_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH))):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.pqo_setup.restype = vp
        L.pqo_setup.argtypes = [C.POINTER(abi.PqpParams), C.c_int, C.c_int, vp, C.c_int, vp]
        L.pqo_free.argtypes = [vp]
        L.pqo_solve.argtypes = [vp]
        L.pqo_update.argtypes = [vp, vp, vp, vp]
        L.pqo_update_full.argtypes = [vp, vp, C.c_int, vp]
        for f in ("pqo_nv", "pqo_m", "pqo_iters", "pqo_status", "pqo_rho_updates", "pqo_nnz_L",
                  "pqo_nnz_A"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = C.c_int
        for f in ("pqo_rho", "pqo_cost", "pqo_pri_res", "pqo_dua_res"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = C.c_double
        for f in ("pqo_get_x", "pqo_get_y", "pqo_get_z"):
            getattr(L, f).argtypes = [vp, vp]
        L.pqo_get_scaled_iterates.argtypes = [vp, vp, vp, vp]
        L.pqo_get_scaling.argtypes = [vp, vp, vp, dp]
        L.pqo_get_sol.argtypes = [vp, vp, C.c_int]
        L.pqo_get_problem.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.pqo_frenet_to_cartesian.argtypes = [C.c_int] + [vp] * 8
        L.pqo_solve_batch.restype = C.c_double
        L.pqo_solve_batch.argtypes = [C.POINTER(abi.PqpParams), C.POINTER(abi.PqpBatchIn),
                                      C.POINTER(abi.PqpBatchOut), C.c_int, C.c_int, C.c_int]
        L.pqo_max_threads.restype = C.c_int
        L.pqo_omp_probe.argtypes = [C.c_int]
        L.pqo_omp_probe.restype = C.c_int
        L.pqo_termination_batch.argtypes = [C.POINTER(abi.PqpParams), C.POINTER(abi.PqpBatchIn), vp, vp, vp, vp, C.c_int]
        L.pqo_batch_setup.restype = vp
        L.pqo_batch_setup.argtypes = [C.POINTER(abi.PqpParams), C.POINTER(abi.PqpBatchIn), C.c_int]
        L.pqo_batch_update_full.argtypes = [vp, C.POINTER(abi.PqpBatchIn), C.c_int]
        L.pqo_batch_solve.restype = C.c_double
        L.pqo_batch_solve.argtypes = [vp, C.POINTER(abi.PqpBatchOut), C.c_int]
        L.pqo_batch_free.argtypes = [vp]
        _lib = L
    return _lib


class OracleSolver:
    """One instance: mirrors BaseSolver (solve, then updateProblemFormulationAndSolve)."""

    def __init__(self, params, knots, inst, n, p=None):
        self.L = lib()
        self.n = int(n)
        self.p = self.n if p is None else int(p)
        self.params = params
        knots = np.ascontiguousarray(knots, dtype=np.float64)
        inst = np.ascontiguousarray(inst, dtype=np.float64)
        self.stride = knots.shape[1]
        self.ws = self.L.pqo_setup(C.byref(params), self.n, self.p, knots.ctypes.data,
                                   self.stride, inst.ctypes.data)
        if not self.ws:
            raise RuntimeError("pqo_setup failed")
        self.nv, self.m = self.L.pqo_nv(self.ws), self.L.pqo_m(self.ws)

    def __del__(self):
        if getattr(self, "ws", None):
            self.L.pqo_free(self.ws)
            self.ws = None

    def solve(self):
        return self.L.pqo_solve(self.ws)

    def update(self, l, psi, k):
        l, psi, k = (np.ascontiguousarray(v[: self.n], dtype=np.float64) for v in (l, psi, k))
        rc = self.L.pqo_update(self.ws, l.ctypes.data, psi.ctypes.data, k.ctypes.data)
        if rc:
            raise RuntimeError("pqo_update failed")

    def update_full(self, knots, inst):
        knots = np.ascontiguousarray(knots, dtype=np.float64)
        inst = np.ascontiguousarray(inst, dtype=np.float64)
        if self.L.pqo_update_full(self.ws, knots.ctypes.data, knots.shape[1], inst.ctypes.data):
            raise RuntimeError("pqo_update_full failed")

    @property
    def iters(self):
        return self.L.pqo_iters(self.ws)

    @property
    def status(self):
        return self.L.pqo_status(self.ws)

    @property
    def rho(self):
        return self.L.pqo_rho(self.ws)

    @property
    def rho_updates(self):
        return self.L.pqo_rho_updates(self.ws)

    @property
    def cost(self):
        return self.L.pqo_cost(self.ws)

    @property
    def residuals(self):
        return self.L.pqo_pri_res(self.ws), self.L.pqo_dua_res(self.ws)

    @property
    def nnz_L(self):
        return self.L.pqo_nnz_L(self.ws)

    def _vec(self, fn, size):
        out = np.zeros(size)
        fn(self.ws, out.ctypes.data)
        return out

    def x(self):
        return self._vec(self.L.pqo_get_x, self.nv)

    def y(self):
        return self._vec(self.L.pqo_get_y, self.m)

    def z(self):
        return self._vec(self.L.pqo_get_z, self.m)

    def scaled_iterates(self):
        x, z, y = np.zeros(self.nv), np.zeros(self.m), np.zeros(self.m)
        self.L.pqo_get_scaled_iterates(self.ws, x.ctypes.data, z.ctypes.data, y.ctypes.data)
        return x, z, y

    def scaling(self):
        D, E, c = np.zeros(self.nv), np.zeros(self.m), C.c_double(0)
        self.L.pqo_get_scaling(self.ws, D.ctypes.data, E.ctypes.data, C.byref(c))
        return D, E, c.value

    def sol(self):
        out = np.zeros((4, self.n))
        self.L.pqo_get_sol(self.ws, out.ctypes.data, self.n)
        return out

    def problem(self):
        """(P diag, A csc, l, u) unscaled, reference index order."""
        nnz = self.L.pqo_nnz_A(self.ws)
        Ap = np.zeros(self.nv + 1, dtype=np.int32)
        Ai = np.zeros(nnz, dtype=np.int32)
        Ax, l, u, Pd = np.zeros(nnz), np.zeros(self.m), np.zeros(self.m), np.zeros(self.nv)
        self.L.pqo_get_problem(self.ws, Ap.ctypes.data, Ai.ctypes.data, Ax.ctypes.data,
                               l.ctypes.data, u.ctypes.data, Pd.ctypes.data)
        A = sp.csc_matrix((Ax, Ai, Ap), shape=(self.m, self.nv))
        return Pd, A, l, u


def solve_batch(params, hb: abi.HostBatch, *, nthreads=1, mode=0, dense_assembly=False,
                full=False):
    """Whole batch through the oracle. Returns (HostResult, wall seconds)."""
    res = abi.HostResult(hb.batch, hb.n_max, full=full)
    bi, bo = hb.as_struct(), res.as_struct()
    secs = lib().pqo_solve_batch(C.byref(params), C.byref(bi), C.byref(bo), int(nthreads),
                                 int(mode), int(bool(dense_assembly)))
    if secs < 0:
        raise RuntimeError("pqo_solve_batch failed")
    return res, secs


def max_threads():
    return lib().pqo_max_threads()


def termination_batch(params, hb: abi.HostBatch, x_full, y_full, z_full, *, nthreads=None):
    """OSQP's unscaled termination test of the given points, FP64, on the oracle-assembled problems
    (linearisation fields of `hb` included). Returns dict of per-instance arrays."""
    x, y, z = (np.ascontiguousarray(v, dtype=np.float64) for v in (x_full, y_full, z_full))
    rep = np.zeros((hb.batch, 6))
    bi = hb.as_struct()
    rc = lib().pqo_termination_batch(C.byref(params), C.byref(bi), x.ctypes.data, y.ctypes.data, z.ctypes.data,
                                     rep.ctypes.data, int(nthreads or max_threads()))
    if rc:
        raise RuntimeError("pqo_termination_batch failed")
    return dict(pri_res=rep[:, 0], eps_pri=rep[:, 1], dua_res=rep[:, 2], eps_dua=rep[:, 3], z_violation=rep[:, 4],
                cost=rep[:, 5])


class OracleBatch:
    """One persistent OSQP-style workspace per path (the reference's solver object), for warm
    re-solve sequences: solve() / update_full(hb) / solve() ..."""

    def __init__(self, params, hb: abi.HostBatch, nthreads=None):
        self.L, self.params = lib(), params
        self.nthreads = int(nthreads or max_threads())
        self.batch, self.n_max = hb.batch, hb.n_max
        bi = hb.as_struct()
        self.pb = self.L.pqo_batch_setup(C.byref(params), C.byref(bi), self.nthreads)
        if not self.pb:
            raise RuntimeError("pqo_batch_setup failed")

    def solve(self, full=False):
        res = abi.HostResult(self.batch, self.n_max, full=full)
        bo = res.as_struct()
        if self.L.pqo_batch_solve(self.pb, C.byref(bo), self.nthreads) < 0:
            raise RuntimeError("pqo_batch_solve failed")
        return res

    def update_full(self, hb: abi.HostBatch):
        bi = hb.as_struct()
        if self.L.pqo_batch_update_full(self.pb, C.byref(bi), self.nthreads):
            raise RuntimeError("pqo_batch_update_full failed")

    def close(self):
        if getattr(self, "pb", None):
            self.L.pqo_batch_free(self.pb)
            self.pb = None

    __del__ = close


def frenet_to_cartesian(ref_xyh, l, psi):
    n = len(l)
    ref_xyh = np.ascontiguousarray(ref_xyh, dtype=np.float64)
    l = np.ascontiguousarray(l, dtype=np.float64)
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    out = np.zeros((3, n))
    lib().pqo_frenet_to_cartesian(n, ref_xyh[0].ctypes.data, ref_xyh[1].ctypes.data,
                                  ref_xyh[2].ctypes.data, l.ctypes.data, psi.ctypes.data,
                                  out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data)
    return out


# --------------------------------------------------------------------------- checkers
def osqp_termination_report(Pd, A, l, u, x, y, z, eps_abs=2e-3, eps_rel=2e-3):
    """OSQP's own unscaled termination test, evaluated in FP64."""
    Ax = A @ x
    Px = Pd * x
    Aty = A.T @ y
    pri_res = np.max(np.abs(Ax - z)) if len(z) else 0.0
    dua_res = np.max(np.abs(Px + Aty))
    eps_pri = eps_abs + eps_rel * max(np.max(np.abs(Ax)), np.max(np.abs(z)))
    eps_dua = eps_abs + eps_rel * max(np.max(np.abs(Px)), np.max(np.abs(Aty)))
    z_in_box = bool(np.all(z >= l - 1e-9) and np.all(z <= u + 1e-9))
    return dict(pri_res=pri_res, dua_res=dua_res, eps_pri=eps_pri, eps_dua=eps_dua,
                ok=bool(pri_res < eps_pri and dua_res < eps_dua), z_in_box=z_in_box)


def kkt_report(Pd, A, l, u, x, y):
    """Independent optimality certificate of (x, y) for min 1/2 x'Px s.t. l <= Ax <= u:
    stationarity, primal feasibility and the sign/complementarity conditions of y."""
    Ax = A @ x
    stat = np.max(np.abs(Pd * x + A.T @ y))
    pfeas = max(np.max(np.maximum(l - Ax, 0)), np.max(np.maximum(Ax - u, 0)))
    gap_lo = Ax - l
    gap_hi = u - Ax
    # y_i > 0 only at the upper bound, y_i < 0 only at the lower bound
    comp = max(np.max(np.maximum(y, 0) * np.minimum(gap_hi, 1e3)),
               np.max(np.maximum(-y, 0) * np.minimum(gap_lo, 1e3)))
    return dict(stationarity=stat, primal_feas=pfeas, complementarity=comp,
                cost=0.5 * float(np.dot(Pd * x, x)))
