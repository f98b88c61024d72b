"""The oracle itself: structure of the assembled QP (closed forms of SURVEY.md §8 / App. A),
KKT optimality of its high-accuracy answers (independent of the ADMM iteration), and the
golden fixture that freezes its behaviour. PARITY UNPINNED: there is no reference-produced
vector for this path (the reference ships no tests and its OSQP dependency is neither vendored
nor installable here), so these checks pin the oracle to the mathematics of the QP."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from path_optimizer_2_b200 import abi, synthetic

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")


@pytest.mark.parametrize("n,p", [(120, None), (240, None), (90, 30), (2, None), (17, 0)])
def test_sizes_and_pattern(n, p):
    k, inst, ne, _ = synthetic.make_instance(3, 0, n)
    s = oracle.OracleSolver(abi.default_params(), k, inst, ne, p)
    pp = n if p is None else p
    assert s.nv == 3 * n + (n - 1) + (pp + n)                 # base_solver.cpp:22-36
    assert s.m == 4 * n + pp + n + 2                           # base_solver.cpp:37
    Pd, A, l, u = s.problem()
    if p is None:
        assert s.nv == 6 * n - 1 and s.m == 6 * n + 2 and A.nnz == 17 * n - 5
    assert np.count_nonzero(Pd) == (n + (n - 1) + (pp + n))   # weight_l = 0 rows are dropped
    # row classes: 3n equalities, kappa boxes at +-tan(35deg)/2.5
    assert np.all(l[:3 * n] == u[:3 * n])
    assert np.allclose(u[3 * n:4 * n], 0.2800830, atol=1e-6) and np.allclose(l[3 * n:4 * n], -0.2800830, atol=1e-6)
    assert l[-2] == -1.0 and u[-2] == 1.0


def test_soft_bounds_rule():
    # base_solver.cpp:290-296 through the assembled bounds
    k, inst, ne, _ = synthetic.make_instance(3, 1, 10)
    k[abi.F_B0_LB, 3], k[abi.F_B0_UB, 3] = -2.0, 3.0      # clearance 5 -> shrink 0.6 each side
    k[abi.F_B1_LB, 3], k[abi.F_B1_UB, 3] = -0.5, 0.5      # clearance 1 -> keep 0.1 -> shrink 0.45
    k[abi.F_B0_LB, 4], k[abi.F_B0_UB, 4] = 0.0, 0.0       # blocked -> equality row
    s = oracle.OracleSolver(abi.default_params(), k, inst, ne)
    _, _, l, u = s.problem()
    r = 4 * 10 + 2 * 3
    assert np.isclose(l[r], -1.4) and np.isclose(u[r], 2.4)
    assert np.isclose(l[r + 1], -0.05) and np.isclose(u[r + 1], 0.05)
    assert l[r + 2] == 0.0 and u[r + 2] == 0.0


@pytest.mark.parametrize("n", [20, 120, 240])
def test_high_accuracy_solution_is_kkt_optimal(n):
    hi = abi.default_params(eps_abs=1e-9, eps_rel=1e-9, max_iter=200000)
    for idx in range(3):
        k, inst, ne, _ = synthetic.make_instance(3, idx, n)
        s = oracle.OracleSolver(hi, k, inst, ne)
        assert s.solve() == abi.PQP_SOLVED
        Pd, A, l, u = s.problem()
        rep = oracle.kkt_report(Pd, A, l, u, s.x(), s.y())
        assert rep["stationarity"] < 1e-6 and rep["primal_feas"] < 1e-6 and rep["complementarity"] < 1e-5


def test_default_tolerance_solution_passes_osqp_test():
    prm = abi.default_params()
    for idx in range(4):
        k, inst, ne, _ = synthetic.make_instance(3, idx, 120)
        s = oracle.OracleSolver(prm, k, inst, ne)
        assert s.solve() == abi.PQP_SOLVED and s.iters % 25 == 0
        Pd, A, l, u = s.problem()
        assert oracle.osqp_termination_report(Pd, A, l, u, s.x(), s.y(), s.z())["ok"]
        # warm re-solve about the first result converges in fewer iterations
        it1 = s.iters
        sol = s.sol()
        s.update(sol[0], sol[1], sol[2])
        assert s.solve() == abi.PQP_SOLVED and s.iters <= it1


def test_infeasible_instance_is_detected():
    k, inst, ne, _ = synthetic.make_instance(103, 4, 3)
    s = oracle.OracleSolver(abi.default_params(), k, inst, ne)
    assert s.solve() == abi.PQP_PRIMAL_INFEASIBLE


def test_batch_driver_matches_single_and_threads():
    prm = abi.default_params()
    hb = synthetic.make_batch(3, 6, 60)
    r1, _ = oracle.solve_batch(prm, hb, nthreads=1, full=True)
    r4, _ = oracle.solve_batch(prm, hb, nthreads=4, full=True, dense_assembly=True)
    assert np.array_equal(r1.x_full, r4.x_full) and np.array_equal(r1.iters, r4.iters)
    for b in range(hb.batch):
        s = oracle.OracleSolver(prm, hb.knots[b], hb.inst[b], int(hb.n[b]))
        s.solve()
        assert np.array_equal(r1.x_full[b, :s.nv], s.x()) and r1.iters[b] == s.iters
    r2, _ = oracle.solve_batch(prm, hb, nthreads=2, mode=1)
    assert np.all(r2.status == abi.PQP_SOLVED) and np.all(r2.iters > r1.iters)


def test_frenet_to_cartesian():
    ref = np.array([[0.0, 1.0], [0.0, 2.0], [0.0, np.pi - 0.1]])
    out = oracle.frenet_to_cartesian(ref, np.array([1.0, -2.0]), np.array([0.2, 0.3]))
    assert np.allclose(out[:, 0], [0.0, 1.0, 0.2])
    h = np.pi - 0.1
    assert np.allclose(out[:, 1], [1.0 - 2.0 * np.cos(h + np.pi / 2 - 2 * np.pi), 2.0 - 2.0 * np.sin(h + np.pi / 2 - 2 * np.pi),
                                   h + 0.3 - 2 * np.pi])


def test_golden_fixture():
    """Frozen oracle outputs (made by tests/golden/make_golden.py): guards the oracle against
    silent drift. These are oracle-produced numbers, not reference-produced ones."""
    with open(GOLDEN) as f:
        gold = json.load(f)
    prm = abi.default_params()
    for g in gold["cases"]:
        k, inst, ne, _ = synthetic.make_instance(g["cfg"], g["index"], g["n"])
        s = oracle.OracleSolver(prm, k, inst, ne)
        assert s.solve() == g["status"] and s.iters == g["iters"]
        assert np.isclose(s.cost, g["cost"], rtol=1e-9, atol=1e-12)
        assert np.allclose(s.sol()[:, ::g["stride"]].ravel(), g["sol_sampled"], rtol=1e-8, atol=1e-10)
