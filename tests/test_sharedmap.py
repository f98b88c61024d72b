"""Shared-obstacle-map workload (BASELINE configs[0..1]): the bounds front end restated in
path_optimizer_2_b200/sharedmap.py, and the solve path on its instances."""
import math

import numpy as np
import pytest

from path_optimizer_2_b200 import abi, sharedmap
from tests import parity


@pytest.fixture(scope="module")
def dmap():
    return sharedmap.DistanceMap()


class _Corridor(sharedmap.DistanceMap):
    """Analytic map: free corridor |y| < half_width along x, distance = half_width - |y|."""

    def __init__(self, half_width):
        self.hw = half_width
        self.lx = self.ly = 200.0

    def lookup(self, x, y):
        return np.maximum(self.hw - np.abs(np.asarray(y, dtype=np.float64)), 0.0) + 0.0 * np.asarray(x)


def test_distance_field_and_lookup(dmap):
    assert dmap.dist.shape == (701, 710)
    assert dmap.dist.min() == 0.0 and 15.0 < dmap.dist.max() < 16.0
    # at cell centres the bilinear lookup returns the cell value; outside the map it is 0
    for i, j in ((10, 20), (350, 355), (699, 708)):
        x, y = dmap.lx / 2 - (i + 0.5) * sharedmap.RES, dmap.ly / 2 - (j + 0.5) * sharedmap.RES
        assert abs(float(dmap.lookup(x, y)) - dmap.dist[i, j]) < 1e-9
    assert float(dmap.lookup(dmap.lx / 2 + 1.0, 0.0)) == 0.0
    # halfway between two cell centres: the mean
    i, j = 350, 355
    x, y = dmap.lx / 2 - (i + 1.0) * sharedmap.RES, dmap.ly / 2 - (j + 0.5) * sharedmap.RES
    assert abs(float(dmap.lookup(x, y)) - 0.5 * (dmap.dist[i, j] + dmap.dist[i + 1, j])) < 1e-9


def test_clearance_in_analytic_corridor():
    # corridor half width 3.0, point on the axis heading +x: the 0.5 m disc first touches the
    # wall at lateral 2.5; coarse step 0.3 stops at 2.7 -> 2.4, fine steps of 0.05 reach 2.5
    # (c < 0.5 is strict, so 2.5 itself is accepted), shrink by 0.5, safety margin 0.3
    cm = _Corridor(3.0)
    ub, lb = sharedmap.clearance(cm, np.array([0.0]), np.array([0.0]), np.array([0.0]))
    assert abs(ub[0] - (2.5 - 0.5 - 0.3)) < 0.051 and abs(lb[0] + ub[0]) < 1e-9
    # off-axis by +1.0: the left wall is 1 m closer. On the right the coarse march stops at 3.6
    # (-> 3.3); the reference's fine pass multiplies the NEGATIVE right bound with the right-hand
    # direction (reference_path_impl.cpp:288-291), i.e. it probes the mirrored point on the left,
    # which is already inside the wall here, so the right bound keeps its coarse value. The
    # restatement follows the reference.
    ub2, lb2 = sharedmap.clearance(cm, np.array([0.0]), np.array([1.0]), np.array([0.0]))
    assert abs(ub2[0] - (1.5 - 0.5 - 0.3)) < 0.051 and abs(lb2[0] + (3.3 - 0.5 - 0.3)) < 1e-6
    # narrow corridor: space 0.4 -> margin (0.4 - 0.2) / 2 = 0.1 each side
    cm = _Corridor(1.2)
    ub, lb = sharedmap.clearance(cm, np.array([0.0]), np.array([0.0]), np.array([0.0]))
    assert 0.0 < ub[0] < 0.16 and abs(lb[0] + ub[0]) < 1e-9
    # closer than the search radius to a wall: blocked -> {0, 0}
    ub, lb = sharedmap.clearance(cm, np.array([0.0]), np.array([0.8]), np.array([0.0]))
    assert ub[0] == 0.0 and lb[0] == 0.0


def test_spline_helpers_on_a_circle():
    r = 20.0
    th = np.linspace(0.0, 1.5, 40)
    sp = sharedmap.SplinePath(r * th, r * np.cos(th), r * np.sin(th))
    s = np.array([5.0, 12.0, 20.0])
    assert np.allclose(sp.curvature(s), 1.0 / r, atol=2e-4)
    assert np.allclose(sp.heading(s), s / r + math.pi / 2, atol=1e-4)
    # projecting a point 1 m outside the circle along the normal returns the foot point
    hh = sp.heading(s)
    tx, ty = sp.xs(s) + 1.0 * np.sin(hh), sp.ys(s) - 1.0 * np.cos(hh)
    px, py = sp.directional_projection(tx, ty, hh + math.pi / 2, s + 0.4)
    assert np.allclose(px, sp.xs(s), atol=1e-4) and np.allclose(py, sp.ys(s), atol=1e-4)
    kn = sharedmap.build_knots(sp, 60)
    assert kn is not None and np.allclose(np.diff(kn[0]), 0.3)  # |k| = 0.05 < 0.08 -> 0.3 m


def test_instances_are_deterministic_and_well_formed(dmap):
    hb, ref = sharedmap.make_batch(6, 60, with_ref=True, dmap=dmap)
    hb2 = sharedmap.make_batch(3, 60, first=3, dmap=dmap)
    assert np.array_equal(hb.knots[3:], hb2.knots) and np.array_equal(hb.inst[3:], hb2.inst)
    k = hb.knots
    assert np.all(k[:, abi.F_B0_UB] - k[:, abi.F_B0_LB] > 0.05)
    assert np.all(k[:, abi.F_B1_UB] - k[:, abi.F_B1_LB] > 0.05)
    assert np.all(np.diff(k[:, abi.F_S], axis=1) > 0.149) and np.all(np.diff(k[:, abi.F_S], axis=1) < 0.301)
    assert np.all(dmap.lookup(ref[:, 0], ref[:, 1]) > 0.5)


def test_oracle_and_emulated_kernel_on_shared_map(dmap):
    from tests.emu import emu
    params = abi.default_params()
    hb = sharedmap.make_batch(4, 120, dmap=dmap)
    es = emu.EmuSolver(params, hb.n_max, hb.batch)
    res = es.solve(hb)
    solved = 0
    for b in range(hb.batch):
        s = parity.oracle_reference(params, hb, b)
        parity.check_instance(params, hb, res, b, oracle_solver=s, label="sharedmap emu")
        solved += int(s.status == abi.PQP_SOLVED)
    assert solved >= 3


@pytest.mark.gpu
def test_gpu_parity_on_shared_map(dmap):
    """BASELINE configs[0] (one path, n=120) and a slice of configs[1] through the C ABI."""
    from path_optimizer_2_b200 import solver
    params = abi.default_params()
    for batch in (1, 48):
        hb, ref = sharedmap.make_batch(batch, 120, with_ref=True, dmap=dmap)
        sv = solver.PathQpSolver(params, n_max=120, batch_max=batch)
        res = sv.solve(hb, full=True)
        for b in range(0, batch, 3):
            s = parity.oracle_reference(params, hb, b)
            parity.check_instance(params, hb, res, b, oracle_solver=s, label="sharedmap gpu")
        ok = res.status == abi.PQP_SOLVED
        assert ok.mean() > 0.9
        # the optimised path stays inside the map's free space (anchors respect the bounds)
        xy = sv.frenet_to_cartesian(hb.n, ref, res.sol)
        assert np.all(dmap.lookup(xy[ok, 0], xy[ok, 1]) > 0.5)
        sv.close()
