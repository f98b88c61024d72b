"""Oracle of the next SURVEY.md §8 row (f-3, TensionSmoother2::osqpSmooth): structure and closed-form
checks of oracle/smoother_oracle.py. Nothing on the GPU consumes it yet; it is here so that the next
widening starts from a pinned oracle."""
import numpy as np

from oracle import osqp_generic, smoother_oracle as so


def _line(p=40, ds=0.5, heading=0.3):
    s = ds * np.arange(p)
    return s * np.cos(heading), s * np.sin(heading), np.full(p, heading), np.zeros(p), s


def test_sizes_and_structure():
    x, y, a, k, s = _line(p=25)
    H, q, A, lo, up = so.assemble(x, y, a, k, s)
    p = 25
    assert H.shape == (4 * p - 1, 4 * p - 1) and A.shape == (3 * (p - 1) + 2, 4 * p - 1)  # tension_smoother_2.cpp:31-32
    assert np.array_equal(lo, up) and abs(H - H.T).max() == 0.0
    assert H.nnz == 2 * p + (p - 1) + 2 * (p - 2)   # x, y, k diagonals + the curvature-rate off-diagonals
    assert A.nnz == 3 * 2 * (p - 1) + 3 * (p - 1) + 2
    assert np.allclose(q[:p], -2 * so.W_DEVIATION * x) and np.all(q[2 * p:] == 0.0)


def test_a_straight_line_is_a_fixed_point():
    x, y, a, k, s = _line()
    ok, rx, ry, rs, g = so.osqp_smooth(x, y, a, k, s)
    assert ok and np.max(np.abs(rx - x)) < 2e-3 and np.max(np.abs(ry - y)) < 2e-3 and np.allclose(rs, s, atol=5e-3)


def test_a_kink_is_smoothed_and_the_start_is_kept():
    p, ds = 60, 0.5
    s = ds * np.arange(p)
    a = np.where(np.arange(p) < p // 2, 0.0, 0.5)      # a 0.5 rad kink in the middle
    x = np.concatenate(([0.0], np.cumsum(ds * np.cos(a[:-1]))))
    y = np.concatenate(([0.0], np.cumsum(ds * np.sin(a[:-1]))))
    k = np.zeros(p)
    k[p // 2 - 1] = 0.5 / ds                           # the raw reference's curvature spike
    ok, rx, ry, rs, g = so.osqp_smooth(x, y, a, k, s)
    assert ok and abs(rx[0]) < 1e-3 and abs(ry[0]) < 1e-3
    xs, ys, zs = g.solution()
    ksm = xs[3 * p:]
    assert np.max(np.abs(ksm)) < 0.25 * (0.5 / ds)     # the spike is spread out ...
    assert np.max(np.abs(np.diff(ksm))) < 0.1          # ... into a slowly varying curvature
    assert np.max(np.hypot(rx - x, ry - y)) < 1.5      # while staying near the raw points
    H, q, A, lo, up = so.assemble(x, y, a, k, s)
    hi = osqp_generic.GenericOsqp(H, q, A, lo, up, eps_abs=1e-9, eps_rel=1e-9, max_iter=50000)
    assert hi.solve() == osqp_generic.SOLVED
    xh, yh, _ = hi.solution()
    assert np.max(np.abs(H @ xh + q + A.T @ yh)) < 1e-6 and np.max(np.abs(A @ xh - lo)) < 1e-6  # KKT of the equality QP
    assert np.max(np.abs(xs[:2 * p] - xh[:2 * p])) < 5e-2  # the eps = 1e-3 answer sits near the exact one
