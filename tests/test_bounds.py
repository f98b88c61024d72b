"""Clearance-bounds front end and the shared-obstacle-map workload
(BASELINE configs[0..1]).

CPU (`-m "not gpu"`): the oracle (oracle/bounds_oracle.py) against closed-form cases, the kernel's
per-thread source compiled for the host (tests/emu/bounds_driver.cpp) against the oracle, and the
emulated solver on shared-map instances. GPU (`-m gpu`): the CUDA kernel through the C ABI against
the oracle, and the map -> bounds -> solve pipeline."""
import math

import numpy as np
import pytest

from oracle import bounds_oracle as bo
from path_optimizer_2_b200 import abi, sharedmap
from tests import parity
from tests.emu import bounds_emu


@pytest.fixture(scope="module")
def dmap():
    return sharedmap.DistanceMap()


@pytest.fixture(scope="module")
def lines(dmap):
    return sharedmap.make_lines(24, 120, dmap=dmap)


def _corridor(half_width):
    """Analytic map: free corridor |y| < half_width along x, distance = half_width - |y|."""
    return lambda x, y: np.maximum(half_width - np.abs(np.asarray(y, dtype=np.float64)), 0.0) + 0.0 * np.asarray(x)


def pinch_map(gap, rows=300, cols=400, res=0.2):
    """Free corridor of half width 4 m along +y... the map's second axis, pinched to `gap` metres
    over 2 m in the middle; returns a sharedmap.DistanceMap built from a synthetic occupancy image."""
    from scipy.ndimage import distance_transform_edt
    free = np.zeros((rows, cols), dtype=bool)
    mid = rows // 2
    free[mid - int(4.0 / res):mid + int(4.0 / res), :] = True
    c0, c1 = cols // 2 - int(1.0 / res), cols // 2 + int(1.0 / res)
    free[:, c0:c1] = False
    half = int(round(0.5 * gap / res))
    free[mid - half:mid + half, c0:c1] = True
    dist = (distance_transform_edt(free).astype(np.float32)) * np.float32(res)
    return sharedmap.DistanceMap(dist=dist, res=res)


def straight_line(dm, n, start_y, length):
    """A reference line along -y... world y decreases with the column index, so the line runs along
    the corridor of `pinch_map` from world y = start_y in the -y direction."""
    sc = np.arange(0.0, length + 1e-9, 1.5)
    xs, ys = np.zeros_like(sc), start_y - sc
    rows = sharedmap.natural_spline_rows(sc, xs, ys)
    st = sharedmap.reference_states(rows, n)
    lb = sharedmap.LineBatch(1, n, rows.shape[1])
    lb.spline[0], lb.k[0] = rows, rows.shape[1]
    lb.states[0] = np.stack(st[:4])
    return lb


# ------------------------------------------------------------------ oracle vs closed forms (CPU)
def test_distance_layer_and_lookups_agree(dmap):
    assert dmap.dist.shape == (701, 710) and dmap.dist.dtype == np.float32
    assert dmap.dist.min() == 0.0 and 15.0 < dmap.dist.max() < 16.0
    for i, j in ((10, 20), (350, 355), (699, 708)):  # at cell centres: the cell value
        x, y = dmap.lx / 2 - (i + 0.5) * dmap.res, dmap.ly / 2 - (j + 0.5) * dmap.res
        assert abs(float(bo.map_distance(dmap.dist, dmap.res, x, y)) - dmap.dist[i, j]) < 1e-6
    assert float(bo.map_distance(dmap.dist, dmap.res, dmap.lx / 2 + 1.0, 0.0)) == 0.0  # Map.cpp:19-21
    i, j = 350, 355  # halfway between two cell centres: the mean
    x, y = dmap.lx / 2 - (i + 1.0) * dmap.res, dmap.ly / 2 - (j + 0.5) * dmap.res
    assert abs(float(bo.map_distance(dmap.dist, dmap.res, x, y)) - 0.5 * (float(dmap.dist[i, j]) + float(dmap.dist[i + 1, j]))) < 1e-6
    rng = np.random.default_rng(0)
    px, py = rng.uniform(-80, 80, 4000), rng.uniform(-80, 80, 4000)
    assert np.array_equal(bo.map_distance(dmap.dist, dmap.res, px, py), dmap.lookup(px, py))


def test_oracle_clearance_in_analytic_corridor():
    # half width 3.0, state on the axis heading +x: the 0.5 m disc touches the wall at lateral 2.5;
    # the coarse march stops at 2.7 (-> 2.4), fine steps of 0.05 reach 2.5 (c < 0.5 is strict),
    # shrink by car_width/2 - 0.5, hard safety margin 0.3
    look = _corridor(3.0)
    lb, ub = bo.clearance(None, None, [0.0], [0.0], [0.0], lookup=look)
    assert abs(ub[0] - (2.5 - 0.5 - 0.3)) < 0.051 and abs(lb[0] + ub[0]) < 1e-9
    # off-axis by +1: the left wall is 1 m closer. On the right the coarse march stops at 3.6 (-> 3.3);
    # the reference's fine pass multiplies the NEGATIVE right bound with the right-hand direction
    # (reference_path_impl.cpp:288-291), i.e. probes the mirrored point on the left, which is inside
    # the wall here, so the right bound keeps its coarse value. The restatement follows the reference.
    lb2, ub2 = bo.clearance(None, None, [0.0], [1.0], [0.0], lookup=look)
    assert abs(ub2[0] - (1.5 - 0.5 - 0.3)) < 0.051 and abs(lb2[0] + (3.3 - 0.5 - 0.3)) < 1e-6
    # narrow corridor: space 0.4 -> margin (0.4 - 0.2) / 2 = 0.1 each side
    lb, ub = bo.clearance(None, None, [0.0], [0.0], [0.0], lookup=_corridor(1.2))
    assert 0.0 < ub[0] < 0.16 and abs(lb[0] + ub[0]) < 1e-9
    # closer than the search radius to a wall -> {0, 0}; narrower than the car -> {0, 0}
    lb, ub = bo.clearance(None, None, [0.0], [0.8], [0.0], lookup=_corridor(1.2))
    assert ub[0] == 0.0 and lb[0] == 0.0
    lb, ub = bo.clearance(None, None, [0.0], [0.0], [0.0], lookup=_corridor(0.9))
    assert ub[0] == 0.0 and lb[0] == 0.0


def test_oracle_spline_against_scipy_and_reference_rules():
    rng = np.random.default_rng(1)
    sx = np.cumsum(rng.uniform(0.5, 2.0, 17))
    xv, yv = np.cumsum(rng.normal(0, 1, 17)), np.cumsum(rng.normal(0, 1, 17))
    rows = sharedmap.natural_spline_rows(sx, xv, yv)  # scipy construction
    for which, v in ((0, xv), (1, yv)):               # the oracle's own tridiagonal solve
        a, b, c = bo.natural_spline(sx, v)
        assert np.allclose(a, rows[1 + 4 * which], atol=1e-10) and np.allclose(b, rows[2 + 4 * which], atol=1e-10)
        assert np.allclose(c, rows[3 + 4 * which], atol=1e-10)
        assert abs(b[0]) < 1e-12 and abs(b[-1]) < 1e-12 and a[-1] == 0.0  # natural ends
    sp = bo.Spline2(rows)
    v, d1, d2 = sp.x(sx)  # at a knot the previous segment is used (lower_bound - 1): same value
    assert np.allclose(v, xv, atol=1e-10)
    # extrapolation: left uses (b0 h + c0) h + y0 and second derivative 2 b0 h; right is the
    # quadratic through the last knot with the end slope (spline.cpp:237-246, 262-268)
    vl, d1l, d2l = sp.x(np.array([sx[0] - 2.0]))
    assert abs(vl[0] - (xv[0] - 2.0 * rows[3][0])) < 1e-10 and abs(d1l[0] - rows[3][0]) < 1e-12 and abs(d2l[0]) < 1e-12
    vr, d1r, d2r = sp.x(np.array([sx[-1] + 2.0]))
    assert abs(vr[0] - (xv[-1] + 2.0 * rows[3][-1])) < 1e-10 and abs(d1r[0] - rows[3][-1]) < 1e-12 and abs(d2r[0]) < 1e-12


def test_oracle_projection_and_states_on_a_circle():
    r = 20.0
    th = np.linspace(0.0, 1.5, 40)
    rows = sharedmap.natural_spline_rows(r * th, r * np.cos(th), r * np.sin(th))
    sp = bo.Spline2(rows)
    s = np.array([5.0, 12.0, 20.0])
    x, dx, ddx = sp.x(s)
    y, dy, ddy = sp.y(s)
    assert np.allclose((dx * ddy - dy * ddx) / np.power(dx * dx + dy * dy, 1.5), 1.0 / r, atol=2e-4)
    hh = np.arctan2(dy, dx)
    assert np.allclose(hh, s / r + math.pi / 2, atol=1e-4)
    # a point 1 m outside the circle, projected along the normal, returns the foot point
    tx, ty = x + 1.0 * np.sin(hh), y - 1.0 * np.cos(hh)
    px, py = bo.directional_projection(sp, tx, ty, hh + math.pi / 2, s + 5.0, s + 0.4)
    assert np.allclose(px, x, atol=1e-4) and np.allclose(py, y, atol=1e-4)
    st = sharedmap.reference_states(rows, 60)
    assert st is not None and np.allclose(np.diff(st[0]), 0.3)  # |k| = 0.05 < 0.08 -> 0.3 m spacing


def test_bounds_golden_fixture(dmap):
    """Frozen oracle outputs (tests/golden/make_bounds_golden.py): guards the bounds oracle, the map
    loader and the spline rows against silent changes. Oracle-produced; parity unpinned."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "bounds_golden.json")) as f:
        cases = json.load(f)["cases"]
    ln = sharedmap.make_lines(3, 120, dmap=dmap)
    for c in cases:
        b = c["line"]
        rows = np.array(c["spline"])
        assert np.allclose(ln.spline_rows(b), rows, rtol=0, atol=1e-12)  # the workload generator is frozen too
        bounds, n_valid = bo.update_bounds(dmap.dist, dmap.res, rows, *ln.states[b])
        assert n_valid == c["n_valid"] and np.allclose(bounds[:, ::8], np.array(c["bounds_sampled"]), rtol=0, atol=1e-12)
        st = bo.build_states(rows, float(rows[0, -1]))
        assert len(st[0]) == c["total_states"]
        assert np.allclose(np.stack(st)[:, ::16], np.array(c["states_sampled"]), rtol=0, atol=1e-12)


# ------------------------------------------------------- the kernel's source on the host (CPU)
def test_kernel_source_on_host_matches_oracle(dmap, lines):
    knots = np.zeros((lines.batch, abi.NFIELDS, lines.n_max))
    eb, env = bounds_emu.compute(dmap.dist, dmap.res, lines.states, lines.n, lines.spline, lines.k, knots=knots)
    for b in range(lines.batch):
        ob, onv = bo.update_bounds(dmap.dist, dmap.res, lines.spline_rows(b), *lines.states[b])
        assert onv == env[b] == lines.n_max
        assert np.max(np.abs(ob - eb[b])) < 1e-9
    assert np.array_equal(knots[:, abi.F_B0_LB:abi.F_B1_UB + 1], eb[:, 0:4])
    w = eb[:, 1] - eb[:, 0]
    assert w.min() > 0.19 and w.max() < 2 * (6.0 - 0.5 - 0.3) + 1e-9  # min_space .. search range


def test_blocked_path_is_cut_like_the_reference():
    dm = pinch_map(gap=1.6)  # narrower than the 2 m car: left < right -> {0, 0} -> blocked
    lb = straight_line(dm, 100, start_y=22.0, length=40.0)
    eb, env = bounds_emu.compute(dm.dist, dm.res, lb.states, lb.n, lb.spline, lb.k)
    ob, onv = bo.update_bounds(dm.dist, dm.res, lb.spline_rows(0), *lb.states[0])
    assert env[0] == onv and 40 < onv < 100  # cut where the FRONT circle (3.9 m ahead) reaches the pinch
    assert np.max(np.abs(ob[:, :onv] - eb[0][:, :onv])) < 1e-9
    s_cut = lb.states[0, 0, onv]
    assert abs((22.0 - s_cut - 3.9) - 1.0) < 1.0  # the front anchor is at the pinch entrance (y = +1)
    assert eb[0, 0, onv] == eb[0, 1, onv]          # {0, 0} + offset on both sides
    wide = pinch_map(gap=3.4)
    lb = straight_line(wide, 100, start_y=22.0, length=40.0)
    eb, env = bounds_emu.compute(wide.dist, wide.res, lb.states, lb.n, lb.spline, lb.k)
    assert env[0] == 100 and (eb[0, 1] - eb[0, 0]).min() < 1.0  # passable, but tight at the pinch


def test_ragged_batch_and_padding(dmap, lines):
    n = lines.n.copy()
    n[::3] = 37
    eb, env = bounds_emu.compute(dmap.dist, dmap.res, lines.states, n, lines.spline, lines.k)
    full, _ = bounds_emu.compute(dmap.dist, dmap.res, lines.states, lines.n, lines.spline, lines.k)
    assert np.array_equal(env, n)
    assert np.array_equal(eb[0, :, :37], full[0, :, :37]) and np.all(eb[0, :, 37:] == 0.0)


def test_reference_states_from_spline(lines):
    """buildReferenceFromSpline: the kernel's source on the host against the oracle
    walk and against the workload generator's own states."""
    max_s = np.array([lines.spline_rows(b)[0, -1] for b in range(lines.batch)])
    st, cv, n, total = bounds_emu.build_states(lines.spline, lines.k, max_s, lines.n_max)
    assert np.all(n == lines.n_max) and np.all(total > lines.n_max)
    assert np.max(np.abs(st - lines.states)) < 1e-9 and np.max(np.abs(cv - lines.kref)) < 1e-9
    for b in range(0, lines.batch, 5):
        os_, ox, oy, oh, ok = bo.build_states(lines.spline_rows(b), max_s[b])
        assert len(os_) == total[b]
        m = lines.n_max
        assert np.max(np.abs(np.stack((os_, ox, oy, oh))[:, :m] - st[b])) < 1e-9 and np.max(np.abs(ok[:m] - cv[b])) < 1e-9
    # a short spline ends before n_max: n = total < n_max, the rest stays zero; fixed step without
    # dynamic segmentation; a spline whose max_s is an exact multiple of the step includes s = max_s
    short = np.array([7.5] * lines.batch)
    st2, cv2, n2, total2 = bounds_emu.build_states(lines.spline, lines.k, short, lines.n_max, dynamic=False)
    assert np.all(n2 == 26) and np.all(total2 == 26) and np.all(st2[:, :, 26:] == 0.0)
    assert np.allclose(st2[0, 0, :26], 0.3 * np.arange(26), atol=1e-12)


def test_lines_are_deterministic(dmap):
    a = sharedmap.make_lines(4, 60, dmap=dmap)
    b = sharedmap.make_lines(2, 60, first=2, dmap=dmap)
    assert np.array_equal(a.states[2:], b.states) and np.array_equal(a.spline[2:, :, :b.k_max], b.spline)
    assert np.all(np.diff(a.states[:, 0], axis=1) > 0.149) and np.all(np.diff(a.states[:, 0], axis=1) < 0.301)
    assert np.all(dmap.lookup(a.states[:, 1], a.states[:, 2]) > 1.0)


def test_emulated_solver_on_shared_map_instances(dmap):
    from tests.emu import emu
    params = abi.default_params()
    ln = sharedmap.make_lines(4, 120, dmap=dmap)
    bnd, nv = bounds_emu.compute(dmap.dist, dmap.res, ln.states, ln.n, ln.spline, ln.k)
    hb = ln.to_host_batch(bnd, nv)
    res = emu.EmuSolver(params, hb.n_max, hb.batch).solve(hb)
    solved = 0
    for b in range(hb.batch):
        s = parity.oracle_reference(params, hb, b)
        parity.check_instance(params, hb, res, b, oracle_solver=s, label="sharedmap emu")
        solved += int(s.status == abi.PQP_SOLVED)
    assert solved >= 3


# ------------------------------------------------------------------------------ GPU (C ABI)
def _assert_bounds_close(gb, ob, label):
    """FP64 on both sides; the only admissible differences are last-bit effects of sin/cos and FMA
    contraction (< 1e-9) and, rarely, a march decision `c < 0.5` landing on the other side of the
    threshold (one coarse or fine step: 0.05 .. 0.3 m)."""
    d = np.abs(gb - ob)
    flips = d > 1e-9
    assert flips.mean() < 2e-3, "%s: %d of %d bounds differ" % (label, flips.sum(), d.size)
    assert np.all(d[flips] < 0.31), label


@pytest.mark.gpu
def test_gpu_bounds_match_oracle(dmap, lines):
    from path_optimizer_2_b200 import bounds
    pbn = bounds.PathBounds(dmap.dist, dmap.res)
    knots = np.zeros((lines.batch, abi.NFIELDS, lines.n_max))
    gb, gnv = pbn.compute(lines.states, lines.n, lines.spline, lines.k, knots=knots)
    assert pbn.last_kernel_ms > 0.0
    for b in range(lines.batch):
        ob, onv = bo.update_bounds(dmap.dist, dmap.res, lines.spline_rows(b), *lines.states[b])
        assert gnv[b] == onv
        _assert_bounds_close(gb[b], ob, "line %d" % b)
    assert np.array_equal(knots[:, abi.F_B0_LB:abi.F_B1_UB + 1], gb[:, 0:4])
    # ragged n, and the host-compiled kernel source agrees too
    n = lines.n.copy()
    n[::3] = 37
    gb2, gnv2 = pbn.compute(lines.states, n, lines.spline, lines.k)
    assert np.array_equal(gnv2, n) and np.array_equal(gb2[0, :, :37], gb[0, :, :37])
    eb, _ = bounds_emu.compute(dmap.dist, dmap.res, lines.states, lines.n, lines.spline, lines.k)
    _assert_bounds_close(gb, eb, "host-compiled source")
    pbn.close()


@pytest.mark.gpu
def test_gpu_reference_states_match_oracle(dmap, lines):
    from path_optimizer_2_b200 import bounds
    pbn = bounds.PathBounds(dmap.dist, dmap.res)
    max_s = np.array([lines.spline_rows(b)[0, -1] for b in range(lines.batch)])
    knots = np.full((lines.batch, abi.NFIELDS, lines.n_max), 7.0)
    st, cv, n, total = pbn.build_states(lines.spline, lines.k, max_s, lines.n_max, knots=knots)
    est, ecv, en, etotal = bounds_emu.build_states(lines.spline, lines.k, max_s, lines.n_max)
    assert np.array_equal(n, en) and np.array_equal(total, etotal)
    assert np.max(np.abs(st - est)) < 1e-9 and np.max(np.abs(cv - ecv)) < 1e-9
    for b in range(0, lines.batch, 5):
        os_, ox, oy, oh, ok = bo.build_states(lines.spline_rows(b), max_s[b])
        assert len(os_) == total[b]
        assert np.max(np.abs(np.stack((os_, ox, oy, oh))[:, :lines.n_max] - st[b])) < 1e-9
    assert np.array_equal(knots[:, abi.F_S], st[:, 0]) and np.array_equal(knots[:, abi.F_KREF], cv)
    assert np.array_equal(knots[:, abi.F_K], cv) and np.all(knots[:, abi.F_L] == 0.0) and np.all(knots[:, abi.F_B0_LB] == 7.0)
    st2, _, n2, total2 = pbn.build_states(lines.spline, lines.k, np.full(lines.batch, 7.5), lines.n_max, dynamic=False)
    assert np.all(n2 == 26) and np.all(total2 == 26) and np.all(st2[:, :, 26:] == 0.0)
    from path_optimizer_2_b200 import solver
    with pytest.raises(solver.PqpError):
        pbn.build_states(lines.spline, lines.k, max_s, lines.n_max, ds_small=0.4, ds_large=0.3)
    pbn.close()


@pytest.mark.gpu
def test_gpu_blocked_path_and_errors():
    from path_optimizer_2_b200 import bounds, solver
    dm = pinch_map(gap=1.6)
    lb = straight_line(dm, 100, start_y=22.0, length=40.0)
    pbn = bounds.PathBounds(dm.dist, dm.res)
    gb, gnv = pbn.compute(lb.states, lb.n, lb.spline, lb.k)
    ob, onv = bo.update_bounds(dm.dist, dm.res, lb.spline_rows(0), *lb.states[0])
    assert gnv[0] == onv and 40 < onv < 100
    _assert_bounds_close(gb[0][:, :onv], ob[:, :onv], "pinch")
    with pytest.raises(solver.PqpError) as ei:  # k_max < 3
        pbn.compute(lb.states, lb.n, lb.spline[:, :, :2], lb.k)
    assert ei.value.code == abi.PQP_E_INVALID
    pbn.close()
    with pytest.raises(solver.PqpError):
        bounds.PathBounds(np.zeros((1, 1), dtype=np.float32), 0.2)


@pytest.mark.gpu
def test_gpu_map_to_path_pipeline(dmap):
    """BASELINE configs[0] (one path) and a slice of configs[1], device-resident from the splines to
    the second solve: buildReferenceFromSpline -> updateBoundsImproved -> BaseSolver::solve ->
    re-linearise -> updateProblemFormulationAndSolve (path_optimizer.cpp:124-161), all through the
    C ABI with device pointers; each stage is checked against its oracle."""
    import torch
    from path_optimizer_2_b200 import bounds, solver
    params = abi.default_params()
    pbn = bounds.PathBounds(dmap.dist, dmap.res)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    for batch in (1, 48):
        n = 120
        ln = sharedmap.make_lines(batch, n, dmap=dmap)
        max_s = np.array([ln.spline_rows(b)[0, -1] for b in range(batch)])
        d_spline, d_k, d_maxs = (torch.from_numpy(v).to(dev) for v in (ln.spline, ln.k, max_s))
        d_inst = torch.from_numpy(ln.inst).to(dev)
        d_states = torch.zeros((batch, 4, n), dtype=torch.float64, device=dev)
        d_curv = torch.zeros((batch, n), dtype=torch.float64, device=dev)
        d_knots = torch.zeros((batch, abi.NFIELDS, n), dtype=torch.float64, device=dev)
        d_bounds = torch.zeros((batch, 6, n), dtype=torch.float64, device=dev)
        d_n, d_nv = (torch.zeros(batch, dtype=torch.int32, device=dev) for _ in range(2))
        d_sol = torch.zeros((batch, 4, n), dtype=torch.float64, device=dev)
        d_sol2 = torch.zeros((batch, 4, n), dtype=torch.float64, device=dev)
        d_cost = torch.zeros(batch, dtype=torch.float64, device=dev)
        d_status, d_iters, d_status2, d_iters2 = (torch.zeros(batch, dtype=torch.int32, device=dev) for _ in range(4))
        sv = solver.PathQpSolver(params, n_max=n, batch_max=batch)
        si = bounds.StatesIn(batch, n, ln.k_max, d_spline.data_ptr(), d_k.data_ptr(), d_maxs.data_ptr(), 0.15, 0.3, 1)
        so = bounds.StatesOut(d_states.data_ptr(), d_curv.data_ptr(), d_n.data_ptr(), None, d_knots.data_ptr())
        bi = bounds.BoundsIn(batch, n, ln.k_max, d_states.data_ptr(), d_n.data_ptr(), d_spline.data_ptr(), d_k.data_ptr())
        bout = bounds.BoundsOut(d_bounds.data_ptr(), d_nv.data_ptr(), d_knots.data_ptr())
        qi = abi.PqpBatchIn(batch, n, d_knots.data_ptr(), d_inst.data_ptr(), d_nv.data_ptr(), None)
        qo = abi.PqpBatchOut(d_sol.data_ptr(), d_cost.data_ptr(), d_status.data_ptr(), d_iters.data_ptr(), None, None, None, None)
        qo2 = abi.PqpBatchOut(d_sol2.data_ptr(), d_cost.data_ptr(), d_status2.data_ptr(), d_iters2.data_ptr(), None, None, None, None)
        pbn.build_states_device(si, so, stream=stream)
        pbn.compute_device(bi, bout, stream=stream)
        sv.solve_device(qi, qo, stream=stream)
        hb1 = None
        torch.cuda.synchronize()
        hb1 = abi.HostBatch(d_knots.cpu().numpy(), ln.inst.copy(), d_nv.cpu().numpy())
        sv.relinearise_device(batch, d_sol.data_ptr(), d_knots.data_ptr(), stream=stream)
        sv.solve_device(qi, qo2, stream=stream, warm=True)
        torch.cuda.synchronize()
        # stage oracles: states and bounds ...
        assert np.array_equal(hb1.n, ln.n) and np.max(np.abs(d_states.cpu().numpy() - ln.states)) < 1e-9
        gb = d_bounds.cpu().numpy()
        for b in range(0, batch, 7):
            ob, onv = bo.update_bounds(dmap.dist, dmap.res, ln.spline_rows(b), *ln.states[b])
            assert onv == hb1.n[b]
            _assert_bounds_close(gb[b], ob, "pipeline line %d" % b)
        assert np.array_equal(hb1.knots[:, abi.F_B0_LB:abi.F_B1_UB + 1], gb[:, 0:4])
        # ... first solve against the oracle, second solve against the oracle re-linearised about it
        sol1, sol2 = d_sol.cpu().numpy(), d_sol2.cpu().numpy()
        st1, st2 = d_status.cpu().numpy(), d_status2.cpu().numpy()
        res_h = sv.solve(hb1, full=True)  # host API on the same inputs: full iterates for the parity check
        assert np.array_equal(res_h.sol, sol1) and np.array_equal(res_h.status, st1)
        hb2 = hb1.with_linearisation(sol1)
        assert np.array_equal(d_knots.cpu().numpy(), hb2.knots)
        res_h2 = sv.resolve(hb2, full=True)
        assert np.array_equal(res_h2.sol, sol2) and np.array_equal(res_h2.status, st2)
        for b in range(0, batch, 3):
            s = parity.oracle_reference(params, hb1, b)
            parity.check_instance(params, hb1, res_h, b, oracle_solver=s, label="sharedmap gpu")
            if s.status == abi.PQP_SOLVED and st1[b] == abi.PQP_SOLVED:
                s2 = parity.oracle_reference(params, hb1, b, warm_from=sol1[b][:3, :int(hb1.n[b])])
                parity.check_instance(params, hb2, res_h2, b, oracle_solver=s2, label="sharedmap gpu warm")
        ok = st2 == abi.PQP_SOLVED
        assert ok.mean() > 0.9
        xy = sv.frenet_to_cartesian(hb1.n, ln.ref_xyh, sol2)  # the optimised path stays in free space
        assert np.all(dmap.lookup(xy[ok, 0], xy[ok, 1]) > 0.5)
        sv.close()
    pbn.close()
