"""ctypes mirror of include/pqp.h (the C ABI of the batched path-QP solver).

Only POD structs and index constants live here; they are shared by the product binding
(`path_optimizer_2_b200.solver`) and by the test-only oracle binding (`oracle/oracle.py`).
"""
import ctypes as C
import math

import numpy as np

# status codes (include/pqp.h)
PQP_SOLVED = 0
PQP_MAX_ITER_REACHED = 1
PQP_PRIMAL_INFEASIBLE = 2
PQP_DUAL_INFEASIBLE = 3
PQP_SOLVED_INACCURATE = 4
PQP_PRIMAL_INFEASIBLE_INACCURATE = 5
PQP_DUAL_INFEASIBLE_INACCURATE = 6
PQP_NUMERICAL_ERROR = 7
PQP_UNSOLVED = 10

PQP_E_INVALID = -1
PQP_E_NO_DEVICE = -2
PQP_E_CUDA = -3
PQP_E_STATE = -4

# knot-block field order
F_S, F_KREF, F_L, F_PSI, F_K, F_B0_LB, F_B0_UB, F_B1_LB, F_B1_UB = range(9)
NFIELDS = 9
# per-instance scalars
I_L0, I_PSI0, I_K0, I_EPSI_LO, I_EPSI_HI = range(5)
NINST = 5
NINFO = 4
INFTY = 1e30


class PqpParams(C.Structure):
    _fields_ = [
        ("front_length", C.c_double),
        ("rear_length", C.c_double),
        ("wheel_base", C.c_double),
        ("max_steering_angle", C.c_double),
        ("expected_safety_margin", C.c_double),
        ("weight_l", C.c_double),
        ("weight_kappa", C.c_double),
        ("weight_dkappa", C.c_double),
        ("weight_slack", C.c_double),
        ("end_l_lb", C.c_double),
        ("end_l_ub", C.c_double),
        ("rho", C.c_double),
        ("sigma", C.c_double),
        ("alpha", C.c_double),
        ("eps_abs", C.c_double),
        ("eps_rel", C.c_double),
        ("eps_prim_inf", C.c_double),
        ("eps_dual_inf", C.c_double),
        ("adaptive_rho_tolerance", C.c_double),
        ("max_iter", C.c_int32),
        ("check_termination", C.c_int32),
        ("scaling", C.c_int32),
        ("adaptive_rho", C.c_int32),
        ("adaptive_rho_interval", C.c_int32),
        ("reserved", C.c_int32),
    ]


def default_params(**overrides) -> PqpParams:
    """Reference flag defaults (planning_flags.cpp:16-22,95), hard-coded weights
    (base_solver.cpp:123-126) and OSQP 0.6.x defaults with eps = 2e-3 (base_solver.cpp:61-62).
    Must agree with pqp_default_params() in the library (tested)."""
    p = PqpParams(
        front_length=3.9, rear_length=-1.0, wheel_base=2.5,
        max_steering_angle=35.0 * math.pi / 180.0, expected_safety_margin=0.6,
        weight_l=0.0, weight_kappa=20.0, weight_dkappa=100.0, weight_slack=10.0,
        end_l_lb=-1.0, end_l_ub=1.0,
        rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=2e-3, eps_rel=2e-3,
        eps_prim_inf=1e-4, eps_dual_inf=1e-4, adaptive_rho_tolerance=5.0,
        max_iter=4000, check_termination=25, scaling=10, adaptive_rho=1,
        adaptive_rho_interval=25, reserved=0)
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class PqpBatchIn(C.Structure):
    _fields_ = [
        ("batch", C.c_int32),
        ("n_max", C.c_int32),
        ("knots", C.c_void_p),
        ("inst", C.c_void_p),
        ("n", C.c_void_p),
        ("p", C.c_void_p),
    ]


class PqpBatchOut(C.Structure):
    _fields_ = [
        ("sol", C.c_void_p),
        ("cost", C.c_void_p),
        ("status", C.c_void_p),
        ("iters", C.c_void_p),
        ("x_full", C.c_void_p),
        ("y_full", C.c_void_p),
        ("z_full", C.c_void_p),
        ("info", C.c_void_p),
    ]


def _ptr(a):
    return None if a is None else a.ctypes.data


class HostBatch:
    """Host-side (numpy) buffers of one batch in the ABI layout."""

    def __init__(self, knots, inst, n, p=None):
        knots = np.ascontiguousarray(knots, dtype=np.float64)
        assert knots.ndim == 3 and knots.shape[1] == NFIELDS, knots.shape
        self.batch, _, self.n_max = knots.shape
        self.knots = knots
        self.inst = np.ascontiguousarray(inst, dtype=np.float64).reshape(self.batch, NINST)
        self.n = np.ascontiguousarray(n, dtype=np.int32).reshape(self.batch)
        self.p = None if p is None else np.ascontiguousarray(p, dtype=np.int32).reshape(self.batch)

    def as_struct(self) -> PqpBatchIn:
        return PqpBatchIn(self.batch, self.n_max, _ptr(self.knots), _ptr(self.inst),
                          _ptr(self.n), _ptr(self.p))

    def slice(self, lo, hi) -> "HostBatch":
        return HostBatch(self.knots[lo:hi], self.inst[lo:hi], self.n[lo:hi],
                         None if self.p is None else self.p[lo:hi])

    def with_linearisation(self, sol) -> "HostBatch":
        """New batch whose L/PSI/K fields are a previous solution (path_optimizer.cpp:153)."""
        k = self.knots.copy()
        k[:, F_L, :] = sol[:, 0, :]
        k[:, F_PSI, :] = sol[:, 1, :]
        k[:, F_K, :] = sol[:, 2, :]
        return HostBatch(k, self.inst, self.n, self.p)

    @property
    def p_eff(self):
        return self.n if self.p is None else self.p


class HostResult:
    """Output buffers of one batch (numpy, ABI layout)."""

    def __init__(self, batch, n_max, full=False, info=True):
        self.batch, self.n_max = batch, n_max
        self.sol = np.zeros((batch, 4, n_max), dtype=np.float64)
        self.cost = np.zeros(batch, dtype=np.float64)
        self.status = np.full(batch, PQP_UNSOLVED, dtype=np.int32)
        self.iters = np.zeros(batch, dtype=np.int32)
        self.x_full = np.zeros((batch, 6 * n_max - 1), dtype=np.float64) if full else None
        self.y_full = np.zeros((batch, 6 * n_max + 2), dtype=np.float64) if full else None
        self.z_full = np.zeros((batch, 6 * n_max + 2), dtype=np.float64) if full else None
        self.info = np.zeros((batch, NINFO), dtype=np.float64) if info else None

    def as_struct(self) -> PqpBatchOut:
        return PqpBatchOut(_ptr(self.sol), _ptr(self.cost), _ptr(self.status), _ptr(self.iters),
                           _ptr(self.x_full), _ptr(self.y_full), _ptr(self.z_full),
                           _ptr(self.info))

    def take(self, idx) -> "HostResult":
        """The instances `idx` as a new (contiguous) result."""
        idx = np.asarray(idx)
        r = HostResult.__new__(HostResult)
        r.batch, r.n_max = int(len(idx)), self.n_max
        for name in ("sol", "cost", "status", "iters", "x_full", "y_full", "z_full", "info"):
            v = getattr(self, name)
            setattr(r, name, None if v is None else np.ascontiguousarray(v[idx]))
        return r


def sizes(n, p=None):
    """(nv, m) of one instance — base_solver.cpp:22-37."""
    p = n if p is None else p
    return 3 * n + (n - 1) + (p + n), 4 * n + p + n + 2
