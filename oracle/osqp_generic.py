"""osqp_generic.py — an independent second restatement of the OSQP 0.6.x algorithm for a general
sparse QP  min 1/2 x'Px + q'x  s.t.  l <= Ax <= u  (numpy / scipy.sparse, FP64).
TEST INFRASTRUCTURE, like everything under oracle/.

Purpose: (1) cross-check of oracle/pqp_oracle.c — that C port is specialised to the path QP
(diagonal P, q = 0, its own elimination order and LDL'); this file shares no code with it, uses a
general P and q and SuperLU for the KKT solves, and must still produce the same status, the same
iteration count and the same x on the path QPs (tests/test_oracle_crosscheck.py). Two independent
restatements agreeing iteration for iteration is the strongest pin available while the reference's
own OSQP cannot be run here (PARITY UNPINNED). (2) Oracle for the next SURVEY.md §8
row (f3: the smoother QPs of tension_smoother_2.cpp have a non-diagonal P and q != 0).

Follows Stellato et al., "OSQP: an operator splitting solver for quadratic programs" (Math. Prog.
Comp. 2020) and OSQP 0.6.x's defaults: Ruiz equilibration with cost scaling (scaling.c), rho vector
with 1e3 rho on equality rows and RHO_MIN on free rows (auxil.c set_rho_vec), relaxed ADMM step
(update_xz_tilde / update_x / update_z / update_y), unscaled termination test and infeasibility
certificates (auxil.c), adaptive rho at a fixed iteration interval (the same stated deviation as the
C port: OSQP's default interval is wall-clock derived).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

RHO_MIN, RHO_MAX, RHO_TOL, RHO_EQ_OVER_RHO_INEQ = 1e-6, 1e6, 1e-4, 1e3
MIN_SCALING, MAX_SCALING, OSQP_INFTY = 1e-4, 1e4, 1e30
SOLVED, MAX_ITER_REACHED, PRIMAL_INFEASIBLE, DUAL_INFEASIBLE, SOLVED_INACCURATE = 0, 1, 2, 3, 4
PRIMAL_INFEASIBLE_INACCURATE, DUAL_INFEASIBLE_INACCURATE = 5, 6


def _limit(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.minimum(v, MAX_SCALING)


class GenericOsqp:
    def __init__(self, P, q, A, l, u, *, rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3,
                 eps_prim_inf=1e-4, eps_dual_inf=1e-4, max_iter=4000, check_termination=25, scaling=10,
                 adaptive_rho=True, adaptive_rho_interval=25, adaptive_rho_tolerance=5.0):
        self.P = sp.csc_matrix(P, dtype=np.float64)
        self.A = sp.csc_matrix(A, dtype=np.float64)
        self.q = np.asarray(q, dtype=np.float64).copy()
        self.l = np.asarray(l, dtype=np.float64).copy()
        self.u = np.asarray(u, dtype=np.float64).copy()
        self.n, self.m = self.P.shape[0], self.A.shape[0]
        self.prm = dict(rho=rho, sigma=sigma, alpha=alpha, eps_abs=eps_abs, eps_rel=eps_rel, eps_prim_inf=eps_prim_inf,
                        eps_dual_inf=eps_dual_inf, max_iter=max_iter, check_termination=check_termination,
                        scaling=scaling, adaptive_rho=adaptive_rho, adaptive_rho_interval=adaptive_rho_interval,
                        adaptive_rho_tolerance=adaptive_rho_tolerance)
        self._scale()
        self.rho = min(max(rho, RHO_MIN), RHO_MAX)
        self._set_rho_vec()
        self._factor()
        self.x, self.z, self.y = np.zeros(self.n), np.zeros(self.m), np.zeros(self.m)
        self.iters, self.status, self.rho_updates = 0, None, 0

    # ---- scaling.c: scale_data
    def _scale(self):
        n, m = self.n, self.m
        P, A, q = self.P.copy(), self.A.copy(), self.q.copy()
        D, E, c = np.ones(n), np.ones(m), 1.0
        for _ in range(self.prm["scaling"]):
            absP, absA = abs(P), abs(A)
            colP = absP.max(axis=0).toarray().ravel() if P.nnz else np.zeros(n)
            colA = absA.max(axis=0).toarray().ravel() if A.nnz else np.zeros(n)
            rowA = absA.max(axis=1).toarray().ravel() if A.nnz else np.zeros(m)
            Dt = 1.0 / np.sqrt(_limit(np.maximum(colP, colA)))
            Et = 1.0 / np.sqrt(_limit(rowA))
            SD, SE = sp.diags(Dt), sp.diags(Et)
            P = (SD @ P @ SD).tocsc()
            A = (SE @ A @ SD).tocsc()
            q = Dt * q
            D *= Dt
            E *= Et
            colP = abs(P).max(axis=0).toarray().ravel() if P.nnz else np.zeros(n)
            ct = max(colP.mean(), float(_limit(np.array([np.max(np.abs(q)) if n else 0.0]))[0]))
            ct = 1.0 / float(_limit(np.array([ct]))[0])
            P = P * ct
            q = q * ct
            c *= ct
        self.Ps, self.As, self.qs = sp.csc_matrix(P), sp.csc_matrix(A), q
        self.D, self.E, self.c = D, E, c
        self.ls, self.us = self.l * E, self.u * E

    def _set_rho_vec(self):
        free = (self.ls < -OSQP_INFTY * MIN_SCALING) & (self.us > OSQP_INFTY * MIN_SCALING)
        eq = ~free & (self.us - self.ls < RHO_TOL)
        self.ctype = np.where(free, -1, np.where(eq, 1, 0))
        self._fill_rho()

    def _fill_rho(self):
        self.rho_vec = np.where(self.ctype == -1, RHO_MIN, np.where(self.ctype == 1, RHO_EQ_OVER_RHO_INEQ * self.rho, self.rho))

    def _factor(self):
        n = self.n
        K = sp.bmat([[self.Ps + self.prm["sigma"] * sp.identity(n), self.As.T],
                     [self.As, -sp.diags(1.0 / self.rho_vec)]], format="csc")
        self.lu = spla.splu(K)

    # ---- auxil.c: residuals and termination, all on unscaled quantities
    def _info(self):
        Dinv, Einv, cinv = 1.0 / self.D, 1.0 / self.E, 1.0 / self.c
        self.Ax = self.As @ self.x
        self.Px = self.Ps @ self.x
        self.Aty = self.As.T @ self.y
        self.rp = self.Ax - self.z
        self.rd = self.Px + self.qs + self.Aty
        self.pri_res = np.max(np.abs(Einv * self.rp)) if self.m else 0.0
        self.dua_res = cinv * np.max(np.abs(Dinv * self.rd))
        self.obj = cinv * (0.5 * self.x @ self.Px + self.qs @ self.x)

    def _tolerances(self, eps_abs, eps_rel):
        Dinv, Einv, cinv = 1.0 / self.D, 1.0 / self.E, 1.0 / self.c
        ep = eps_abs + eps_rel * max(np.max(np.abs(Einv * self.z), initial=0.0), np.max(np.abs(Einv * self.Ax), initial=0.0))
        ed = eps_abs + eps_rel * cinv * max(np.max(np.abs(Dinv * self.qs)), np.max(np.abs(Dinv * self.Aty)),
                                            np.max(np.abs(Dinv * self.Px)))
        return ep, ed

    def _primal_infeasible(self, eps):
        dy = self.delta_y.copy()
        uinf, linf = self.us > OSQP_INFTY * MIN_SCALING, self.ls < -OSQP_INFTY * MIN_SCALING
        dy = np.where(uinf & linf, 0.0, np.where(uinf, np.minimum(dy, 0.0), np.where(linf, np.maximum(dy, 0.0), dy)))
        norm_dy = np.max(np.abs(self.E * dy), initial=0.0)
        if norm_dy > eps:
            lhs = np.sum(np.where(uinf, 0.0, self.us) * np.maximum(dy, 0.0) + np.where(linf, 0.0, self.ls) * np.minimum(dy, 0.0))
            if lhs < -eps * norm_dy:
                return np.max(np.abs((self.As.T @ dy) / self.D)) < eps * norm_dy
        return False

    def _dual_infeasible(self, eps):
        dx = self.delta_x
        norm_dx = np.max(np.abs(self.D * dx), initial=0.0)
        if norm_dx > eps:
            if self.qs @ dx < -self.c * eps * norm_dx:
                if np.max(np.abs((self.Ps @ dx) / self.D)) < self.c * eps * norm_dx:
                    Adx = (self.As @ dx) / self.E
                    uinf, linf = self.us > OSQP_INFTY * MIN_SCALING, self.ls < -OSQP_INFTY * MIN_SCALING
                    ok_u = np.where(uinf, True, Adx < eps * norm_dx)
                    ok_l = np.where(linf, True, Adx > -eps * norm_dx)
                    return bool(np.all(ok_u & ok_l))
        return False

    def _check(self, approximate):
        p = self.prm
        f = 10.0 if approximate else 1.0
        ep, ed = self._tolerances(f * p["eps_abs"], f * p["eps_rel"])
        prim_ok, dual_ok = self.pri_res < ep, self.dua_res < ed
        prim_inf = (not prim_ok) and self._primal_infeasible(f * p["eps_prim_inf"])
        dual_inf = (not dual_ok) and self._dual_infeasible(f * p["eps_dual_inf"])
        if prim_ok and dual_ok:
            self.status = SOLVED_INACCURATE if approximate else SOLVED
        elif prim_inf:
            self.status = PRIMAL_INFEASIBLE_INACCURATE if approximate else PRIMAL_INFEASIBLE
        elif dual_inf:
            self.status = DUAL_INFEASIBLE_INACCURATE if approximate else DUAL_INFEASIBLE
        else:
            return False
        return True

    def _rho_estimate(self):
        pri = np.max(np.abs(self.rp), initial=0.0) / (max(np.max(np.abs(self.z), initial=0.0), np.max(np.abs(self.Ax), initial=0.0)) + 1e-10)
        dua = np.max(np.abs(self.rd)) / (max(np.max(np.abs(self.qs)), np.max(np.abs(self.Aty)), np.max(np.abs(self.Px))) + 1e-10)
        return min(max(self.rho * np.sqrt(pri / (dua + 1e-10)), RHO_MIN), RHO_MAX)

    # ---- osqp_solve
    def solve(self):
        p = self.prm
        n, alpha, sigma = self.n, p["alpha"], p["sigma"]
        self.status, self.rho_updates = None, 0
        can_check = False
        it = 0
        for it in range(1, p["max_iter"] + 1):
            x_prev, z_prev = self.x, self.z
            rhs = np.concatenate((sigma * x_prev - self.qs, z_prev - self.y / self.rho_vec))
            sol = self.lu.solve(rhs)
            xt = sol[:n]
            zt = rhs[n:] + sol[n:] / self.rho_vec
            self.x = alpha * xt + (1.0 - alpha) * x_prev
            self.delta_x = self.x - x_prev
            zh = alpha * zt + (1.0 - alpha) * z_prev
            self.z = np.minimum(np.maximum(zh + self.y / self.rho_vec, self.ls), self.us)
            self.delta_y = self.rho_vec * (zh - self.z)
            self.y = self.y + self.delta_y
            can_check = p["check_termination"] and it % p["check_termination"] == 0
            if can_check:
                self._info()
                if self._check(False):
                    break
            if p["adaptive_rho"] and p["adaptive_rho_interval"] and it % p["adaptive_rho_interval"] == 0:
                if not can_check:
                    self._info()
                est = self._rho_estimate()
                if est > self.rho * p["adaptive_rho_tolerance"] or est < self.rho / p["adaptive_rho_tolerance"]:
                    self.rho = est
                    self._fill_rho()
                    self._factor()
                    self.rho_updates += 1
        self.iters = it
        if not can_check:
            self._info()
            self._check(False)
        if self.status is None:
            if not self._check(True):
                self.status = MAX_ITER_REACHED
        return self.status

    # unscaled solution
    def solution(self):
        return self.D * self.x, self.E * self.y / self.c, self.z / self.E
