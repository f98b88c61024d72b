"""smoother_oracle.py — CPU restatement of TensionSmoother2::osqpSmooth (not yet
built on the GPU): assembly of the smoother QP as /root/reference/src/reference_path_smoother/
tension_smoother_2.cpp:20-158 does it, solved with the generic OSQP restatement (OSQP defaults,
eps = 1e-3: the reference sets only verbosity and warm start, :33-34). TEST INFRASTRUCTURE;
PARITY UNPINNED (no reference-produced vectors; pinned by closed forms in tests/test_smoother_oracle.py).

Variables (4p - 1): x_0..x_{p-1}, y_0.., theta_0.., k_0..k_{p-2}; constraints (3(p-1) + 2), all equalities:
  x_{i+1} - x_i + ds_i sin(a_i) theta_i  = ds_i cos(a_i)          (:113-121,132-134)
  y_{i+1} - y_i - ds_i cos(a_i) theta_i  = ds_i sin(a_i)
  theta_{i+1} - theta_i - ds_i k_i       = -ds_i k_list_i
  x_0 = x_list_0, y_0 = y_list_0                                   (:123,140-141)
Cost: w_dev sum (x_i - x_list_i)^2 + (y_i - y_list_i)^2 + w_k sum k_i^2 + w_dk sum (k_{i+1} - k_i)^2
(Hessian :76-96 carries the factor 2, gradient :143-157 is -2 w_dev x_list).
"""
import numpy as np
import scipy.sparse as sp

from . import osqp_generic

W_DEVIATION, W_CURVATURE, W_CURVATURE_RATE = 0.005, 1.0, 10.0  # planning_flags.cpp:57-61


def assemble(x_list, y_list, angle_list, k_list, s_list, w_dev=W_DEVIATION, w_k=W_CURVATURE, w_dk=W_CURVATURE_RATE):
    x_list, y_list, angle_list, k_list, s_list = (np.asarray(v, dtype=np.float64) for v in (x_list, y_list, angle_list, k_list, s_list))
    p = len(x_list)
    nv, m = 4 * p - 1, 3 * (p - 1) + 2
    xs, ys, ts, ks = 0, p, 2 * p, 3 * p
    H = sp.lil_matrix((nv, nv))
    for i in range(p):
        H[xs + i, xs + i] = H[ys + i, ys + i] = 2.0 * w_dev
        if i != p - 1:
            H[ks + i, ks + i] = 2.0 * w_k
    for i in range(p - 2):  # 2 w_dk [1 -1; -1 1] on (k_i, k_{i+1})
        H[ks + i, ks + i] += 2.0 * w_dk
        H[ks + i + 1, ks + i + 1] += 2.0 * w_dk
        H[ks + i, ks + i + 1] -= 2.0 * w_dk
        H[ks + i + 1, ks + i] -= 2.0 * w_dk
    q = np.zeros(nv)
    q[xs:xs + p] = -2.0 * w_dev * x_list
    q[ys:ys + p] = -2.0 * w_dev * y_list
    A = sp.lil_matrix((m, nv))
    lo = np.zeros(m)
    cx, cy, ct = 0, p - 1, 2 * (p - 1)
    for i in range(p - 1):
        ds = s_list[i + 1] - s_list[i]
        A[cx + i, xs + i + 1] = A[cy + i, ys + i + 1] = A[ct + i, ts + i + 1] = 1.0
        A[cx + i, xs + i] = A[cy + i, ys + i] = A[ct + i, ts + i] = -1.0
        A[cx + i, ts + i] = ds * np.sin(angle_list[i])
        A[cy + i, ts + i] = -ds * np.cos(angle_list[i])
        A[ct + i, ks + i] = -ds
        lo[cx + i] = ds * np.cos(angle_list[i])
        lo[cy + i] = ds * np.sin(angle_list[i])
        lo[ct + i] = -ds * k_list[i]
    A[3 * (p - 1), xs] = 1.0
    A[3 * (p - 1) + 1, ys] = 1.0
    lo[3 * (p - 1)], lo[3 * (p - 1) + 1] = x_list[0], y_list[0]
    return sp.csc_matrix(H), q, sp.csc_matrix(A), lo, lo.copy()


def osqp_smooth(x_list, y_list, angle_list, k_list, s_list, **weights):
    """-> (ok, result_x, result_y, result_s, solver): OSQP defaults, result_s re-accumulated from the
    smoothed points (:58-70)."""
    H, q, A, lo, up = assemble(x_list, y_list, angle_list, k_list, s_list, **weights)
    g = osqp_generic.GenericOsqp(H, q, A, lo, up)  # eps_abs = eps_rel = 1e-3, max_iter 4000, adaptive rho
    ok = g.solve() == osqp_generic.SOLVED
    x, _, _ = g.solution()
    p = len(x_list)
    rx, ry = x[:p], x[p:2 * p]
    rs = np.concatenate(([0.0], np.cumsum(np.hypot(np.diff(rx), np.diff(ry)))))
    return ok, rx, ry, rs, g


# ----------------------------------------------------------------------------------------------------------------------
# ReferencePathSmoother::postSmooth's QP (/root/reference/src/reference_path_smoother/reference_path_smoother.cpp:526-636)
# Variables (3p): x_0..x_{p-1} (lateral offsets of the DP layers), dx_0.., ddx_0..; constraints (3p - 2):
#   x_i in [layers_bounds_[i]]  (row 0: x_0 = vehicle_l_wrt_smoothed_ref_, :628-633)
#   x_{i+1} - x_i - ds_i dx_i = 0,  dx_{i+1} - dx_i - ds_i ddx_i = 0          (:612-622)
# Cost: 1/2 (1 x^2 + 100 dx^2 + 1000 ddx^2) (Hessian :584-597 = diag(weight), no factor 2), q = 0. OSQP defaults.
POST_W_X, POST_W_DX, POST_W_DDX = 1.0, 100.0, 1000.0


def assemble_post(layer_s, lower, upper, vehicle_l, w_x=POST_W_X, w_dx=POST_W_DX, w_ddx=POST_W_DDX):
    layer_s, lower, upper = (np.asarray(v, dtype=np.float64) for v in (layer_s, lower, upper))
    p = len(layer_s)
    nv, m = 3 * p, 3 * p - 2
    H = sp.diags(np.concatenate((np.full(p, w_x), np.full(p, w_dx), np.full(p, w_ddx)))).tocsc()
    A = sp.lil_matrix((m, nv))
    lo, up = np.zeros(m), np.zeros(m)
    cdx, cddx = p, 2 * p - 1
    for i in range(p):
        A[i, i] = 1.0
    for i in range(p - 1):
        ds = layer_s[i + 1] - layer_s[i]
        A[cdx + i, i + 1], A[cdx + i, i], A[cdx + i, p + i] = 1.0, -1.0, -ds
        A[cddx + i, p + i + 1], A[cddx + i, p + i], A[cddx + i, 2 * p + i] = 1.0, -1.0, -ds
    lo[0] = up[0] = vehicle_l
    lo[1:p], up[1:p] = lower[1:], upper[1:]
    return H, np.zeros(nv), sp.csc_matrix(A), lo, up


def post_smooth(layer_s, lower, upper, vehicle_l, **kw):
    """-> (ok, offsets = QPSolution(0..p-1), solver)."""
    H, q, A, lo, up = assemble_post(layer_s, lower, upper, vehicle_l)
    g = osqp_generic.GenericOsqp(H, q, A, lo, up, **kw)
    ok = g.solve() == osqp_generic.SOLVED
    x, _, _ = g.solution()
    return ok, x[:len(layer_s)], g
