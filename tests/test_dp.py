"""Lattice DP search: oracle closed forms, the kernel's source on the CPU against
the oracle, and the CUDA kernel against the oracle (`-m gpu`).

Parity bar: every index output (feasibility table, parent table, chosen lateral index per layer, number of
layers kept) is EXACT; the cost table and the bounds are FP64 and must agree to 1e-9 relative (costs are sums
of atan2 / sin / cos terms whose libm and CUDA implementations differ by an ulp or two; the strict
comparisons they feed are only exposed to that when two candidates tie within ~1e-15 - the parent tables of
every test path match exactly, so no such tie occurs in them)."""
import numpy as np
import pytest

from oracle import bounds_oracle as bo, dp_oracle
from path_optimizer_2_b200 import sharedmap

DBL_MAX = np.finfo(np.float64).max


def _corridor_map(rows=400, cols=400, res=0.2, half_width=None, wall_y=None):
    """Distance layer of a straight corridor along x (or a single wall at y = wall_y)."""
    xs = 0.5 * rows * res - (np.arange(rows) + 0.5) * res
    ys = 0.5 * cols * res - (np.arange(cols) + 0.5) * res
    Y = np.broadcast_to(ys[None, :], (rows, cols))
    if half_width is not None:
        d = np.maximum(half_width - np.abs(Y), 0.0)
    else:
        d = np.maximum(wall_y - Y, 0.0)
    return d.astype(np.float32), res


def _straight_spline(x0=-30.0, length=40.0, k=21):
    s = np.linspace(0.0, length, k)
    rows = np.zeros((9, k))
    rows[0] = s
    xa, xb, xc = bo.natural_spline(s, x0 + s)
    ya, yb, yc = bo.natural_spline(s, np.zeros(k))
    rows[1:5] = xa, xb, xc, x0 + s
    rows[5:9] = ya, yb, yc, np.zeros(k)
    return rows, length


def test_oracle_straight_corridor_stays_on_the_centre_line():
    dist, res = _corridor_map(half_width=6.0)
    rows, length = _straight_spline()
    r = dp_oracle.graph_search_dp(dist, res, rows, length, (-30.0, 0.0, 0.0))
    assert r["ok"] and r["n_out"] == r["n_layers"] == 28  # 0, 1.5, ..., 39, 40
    ls = np.array(dp_oracle.lateral_offsets())
    assert len(ls) == 34 and abs(ls[-1] - 9.8) < 1e-9
    centre = int(np.argmin(np.abs(ls)))
    # the vehicle starts at l = 0 -> index int(10 / 0.6) = 16 (l = -0.4); offset and angle costs pull towards l ~ 0
    assert r["chosen"][0] == 16 and set(r["chosen"][3:].tolist()) <= {centre - 1, centre, centre + 1}
    # corridor half width 6 m, search threshold 1.2 m: the bounds stop about 4.8 m from the centre line
    assert np.all(np.abs(r["upper"][1:] - 4.8) <= 0.41) and np.all(np.abs(r["lower"][1:] + 4.8) <= 0.41)
    assert r["lower"][0] == -10 and r["upper"][0] == 10
    assert abs(r["vehicle_l"]) < 1e-12 and r["target_s"] == length


def test_oracle_wall_pushes_the_corridor_away_and_far_vehicle_is_refused():
    dist, res = _corridor_map(wall_y=2.0)   # free space below y = 2
    rows, length = _straight_spline()
    r = dp_oracle.graph_search_dp(dist, res, rows, length, (-30.0, -1.0, 0.0))
    assert r["ok"]
    assert np.all(r["upper"][1:] <= 2.0 - 1.2 + 0.21)      # upper bound keeps the 1.2 m threshold from the wall
    assert np.all(r["lower"][1:] <= -5.9)                    # nothing on the other side within the 6 m check limit
    far = dp_oracle.graph_search_dp(dist, res, rows, length, (-30.0, -10.5, 0.0))
    assert not far["ok"]


def _map_batch(batch, n=120, first=0):
    dmap = sharedmap.DistanceMap()
    lines = sharedmap.make_lines(batch, n, first=first, dmap=dmap)
    length = np.array([lines.spline_rows(b)[0, lines.k[b] - 1] + 3.0 for b in range(batch)])  # TensionSmoother adds 3 m
    rng = np.random.default_rng(7 + first)
    start = np.zeros((batch, 3))
    for b in range(batch):
        sp = bo.Spline2(lines.spline_rows(b))
        s0 = rng.uniform(0.0, 2.0)
        x, y = dp_oracle._xy(sp, s0)
        h = dp_oracle.heading(sp, s0)
        off = rng.uniform(-1.0, 1.0)
        start[b] = x - off * np.sin(h), y + off * np.cos(h), h + rng.uniform(-0.2, 0.2)
    return dmap, lines, length, start


def _compare(r, b, o, J):
    tag = "path %d" % b
    if not o["ok"]:
        assert r.status[b] == 0, tag
        return
    assert r.status[b] == 1 and r.n_layers[b] == o["n_layers"] and r.n_out[b] == o["n_out"], tag
    L, n = o["n_layers"], o["n_out"]
    assert np.array_equal(r.feasible[b, :L].astype(bool), o["feasible"]), tag + ": feasibility table"
    reached = o["max_layer"] + 1
    assert np.array_equal(r.parent[b, :reached], o["parent"][:reached]), tag + ": parent table"
    assert np.array_equal(r.chosen[b, :n], o["chosen"]), tag + ": chosen indices"
    co, cg = o["cost"][:reached], r.cost[b, :reached]
    fin = co < DBL_MAX
    assert np.array_equal(fin, cg < DBL_MAX), tag
    assert np.allclose(cg[fin], co[fin], rtol=1e-9, atol=1e-12), tag + ": cost table"
    assert np.allclose(r.layer_s[b, :n], o["layer_s"], rtol=0, atol=1e-9)
    assert np.allclose(r.lower[b, :n], o["lower"], rtol=0, atol=1e-9) and np.allclose(r.upper[b, :n], o["upper"], rtol=0, atol=1e-9)
    assert abs(r.vehicle_l[b] - o["vehicle_l"]) < 1e-9 and r.target_s[b] == o["target_s"]


def test_kernel_source_on_cpu_matches_oracle():
    from tests.emu import dp_emu
    dmap, lines, length, start = _map_batch(12)
    r = dp_emu.search(dmap.dist, dmap.res, lines.spline, lines.k, length, start)
    for b in range(12):
        o = dp_oracle.graph_search_dp(dmap.dist, dmap.res, lines.spline_rows(b), length[b], start[b])
        _compare(r, b, o, 34)
    # synthetic cases incl. the refused one
    dist, res = _corridor_map(wall_y=2.0)
    rows, ln = _straight_spline()
    sp = np.stack([rows, rows])
    st = np.array([[-30.0, -1.0, 0.0], [-30.0, -10.5, 0.0]])
    r = dp_emu.search(dist, res, sp, [21, 21], [ln, ln], st)
    for b in range(2):
        _compare(r, b, dp_oracle.graph_search_dp(dist, res, rows, ln, st[b]), 34)


@pytest.mark.gpu
def test_gpu_dp_matches_oracle():
    from path_optimizer_2_b200 import bounds, dp
    B = 96
    dmap, lines, length, start = _map_batch(B)
    pbn = bounds.PathBounds(dmap.dist, dmap.res, device=0)
    ds = dp.DpSearch(pbn, layers_max=160, batch_max=B)
    assert ds.lateral == 34
    r = ds.search(lines.spline, lines.k, length, start)
    for b in range(B):
        o = dp_oracle.graph_search_dp(dmap.dist, dmap.res, lines.spline_rows(b), length[b], start[b])
        _compare(r, b, o, 34)
    print("dp kernel: %d paths, %.3f ms" % (B, ds.last_kernel_ms))
    # capacity error is per path, not a crash
    small = dp.DpSearch(pbn, layers_max=8, batch_max=4)
    rs = small.search(lines.spline[:4], lines.k[:4], length[:4], start[:4], tables=False)
    assert np.all(rs.status == dp.DP_TOO_MANY_LAYERS)
    small.close()
    ds.close()
    pbn.close()


def test_dp_symbols_declared_and_exported():
    import os
    import re
    from path_optimizer_2_b200 import dp
    L = dp._lib()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "include", "pqp_dp.h")) as f:
        declared = set(m.group(1) for m in re.finditer(r"^\s*(?:int|void|int32_t|const char \*)\s*(pqp_dp_\w+)\s*\(", f.read(), re.M))
    assert declared == set(dp.EXPORTED_SYMBOLS)
    assert all(hasattr(L, s) for s in declared)
