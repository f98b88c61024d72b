"""Cross-check of the C oracle (oracle/pqp_oracle.c, specialised to the path QP) against an
independently written generic OSQP restatement (oracle/osqp_generic.py: general sparse P, q, SuperLU):
no shared code, same published algorithm -> same status, same iteration count, same rho and the
same solution on the path QPs. With the reference's own OSQP unavailable (parity unpinned) this is
the strongest pin of the oracle that can be had here."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle, osqp_generic
from path_optimizer_2_b200 import abi, synthetic


def _generic(s, params):
    Pd, A, l, u = s.problem()
    g = osqp_generic.GenericOsqp(sp.diags(Pd), np.zeros(len(Pd)), A, l, u, rho=params.rho, sigma=params.sigma,
                                 alpha=params.alpha, eps_abs=params.eps_abs, eps_rel=params.eps_rel,
                                 eps_prim_inf=params.eps_prim_inf, eps_dual_inf=params.eps_dual_inf,
                                 max_iter=params.max_iter, check_termination=params.check_termination,
                                 scaling=params.scaling, adaptive_rho=bool(params.adaptive_rho),
                                 adaptive_rho_interval=params.adaptive_rho_interval,
                                 adaptive_rho_tolerance=params.adaptive_rho_tolerance)
    return g


@pytest.mark.parametrize("cfg,n,count", [(3, 120, 6), (3, 240, 3), (7, 60, 6), (103, 3, 6), (100 + 31, 31, 4)])
def test_c_oracle_and_generic_osqp_agree(cfg, n, count):
    params = abi.default_params()
    hb = synthetic.make_batch(cfg, count, n, ragged=(cfg == 7))
    for b in range(count):
        s = oracle.OracleSolver(params, hb.knots[b], hb.inst[b], int(hb.n[b]))
        st = s.solve()
        g = _generic(s, params)
        gst = g.solve()
        assert gst == st and g.iters == s.iters and g.rho_updates == s.rho_updates, (cfg, b, gst, st, g.iters, s.iters)
        assert abs(g.rho - s.rho) <= 1e-7 * s.rho
        if st == abi.PQP_SOLVED:
            x, y, z = g.solution()
            assert np.max(np.abs(x - s.x())) < 1e-7 and np.max(np.abs(z - s.z())) < 1e-7
            assert np.max(np.abs(y - s.y())) < 1e-5 * max(1.0, np.max(np.abs(s.y())))
            assert abs(g.obj - s.cost) < 1e-8 * max(1.0, abs(s.cost))


def test_generic_osqp_on_a_qp_with_linear_term_and_coupled_hessian():
    """The features the path QP does not exercise (q != 0, off-diagonal P - the smoother QPs of the
    next SURVEY.md §8 row): KKT conditions of the returned point at a tight tolerance."""
    rng = np.random.default_rng(5)
    n, m = 30, 40
    M = sp.random(n, n, density=0.15, random_state=7, format="csc")
    P = (M.T @ M + 0.05 * sp.identity(n)).tocsc()
    q = rng.normal(size=n)
    A = sp.vstack([sp.random(m - n, n, density=0.3, random_state=9, format="csc"), sp.identity(n)]).tocsc()
    l = np.concatenate((-np.ones(m - n), -0.5 * np.ones(n)))
    u = np.concatenate((np.ones(m - n), 0.5 * np.ones(n)))
    l[:3] = u[:3] = 0.1  # three equality rows
    g = osqp_generic.GenericOsqp(P, q, A, l, u, eps_abs=1e-9, eps_rel=1e-9, max_iter=20000)
    assert g.solve() == osqp_generic.SOLVED
    x, y, z = g.solution()
    assert np.max(np.abs(P @ x + q + A.T @ y)) < 1e-6 and np.max(np.abs(A @ x - z)) < 1e-6
    assert np.all(z >= l - 1e-9) and np.all(z <= u + 1e-9)
    assert np.all(y[z < u - 1e-6] <= 1e-6) and np.all(y[z > l + 1e-6] >= -1e-6)  # complementarity
