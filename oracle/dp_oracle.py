"""dp_oracle.py — CPU restatement (numpy scalars, FP64) of the reference's lattice DP search
(SURVEY.md §8 row f-4). TEST INFRASTRUCTURE: imported only by tests/ and bench.py's checker legs;
the product path is the CUDA kernel in path_optimizer_2_b200/csrc/pqp_dp.cu and never calls this.

PARITY UNPINNED: the reference has no test or golden vector for this path and cannot be built here
(grid_map, glog, gflags are absent); pinned by closed-form cases in tests/test_dp.py (straight corridor:
the search stays on the centre line; a wall on one side: the path and its bounds move away from it).

Follows, relative to /root/reference/:
  graph_search_dp      src/reference_path_smoother/reference_path_smoother.cpp:142-295 (graphSearchDp)
  node cost            :107-140 (calculateCostAt)
  get_projection       src/tools/tools.cpp:66-128 (getProjection, getProjectionByNewton)
  heading / curvature  src/tools/tools.cpp:32-44
  global2Local         src/tools/tools.cpp:56-63
  map lookups          src/tools/Map.cpp:16-26 through oracle/bounds_oracle.map_distance
  flag defaults        src/config/planning_flags.cpp:10 (car_width 2.0), :38-42 (lateral range 10, longitudinal
                       spacing 1.5, lateral spacing 0.6)
Quirk kept: DpPoint::dir of a never-reached predecessor is uninitialised in the reference; such a
predecessor can never be selected (its cost is DBL_MAX and costs are non-negative), so it is skipped.
"""
import math

import numpy as np

from . import bounds_oracle as bo

LATERAL_RANGE, LON_SPACING, LAT_SPACING, CAR_WIDTH = 10.0, 1.5, 0.6, 2.0
DBL_MAX = float(np.finfo(np.float64).max)
W_REF_OFFSET, W_OBSTACLE, W_ANGLE_CHANGE, W_REF_ANGLE_DIFF, SAFE_DISTANCE = 1.0, 0.5, 16.0, 0.5, 3.0
CHECK_S, CHECK_LIMIT = 0.2, 6.0


def constrain_angle(a):  # include/tools/tools.hpp:24-35
    while a > math.pi:
        a -= 2 * math.pi
    while a < -math.pi:
        a += 2 * math.pi
    return a


def _xy(sp, s):
    return float(sp.x(s)[0]), float(sp.y(s)[0])


def heading(sp, s):
    return math.atan2(float(sp.y(s)[1]), float(sp.x(s)[1]))


def curvature(sp, s):
    _, dx, ddx = (float(v) for v in sp.x(s))
    _, dy, ddy = (float(v) for v in sp.y(s))
    return (dx * ddy - dy * ddx) / math.pow(math.pow(dx, 2) + math.pow(dy, 2), 1.5)


def projection_newton(sp, tx, ty, max_s, hint_s):
    hint_s = min(hint_s, max_s)
    cur, prev = hint_s, hint_s
    for _ in range(20):
        x, dx, ddx = (float(v) for v in sp.x(cur))
        y, dy, ddy = (float(v) for v in sp.y(cur))
        j = (x - tx) * dx + (y - ty) * dy
        h = dx * dx + (x - tx) * ddx + dy * dy + (y - ty) * ddy
        cur -= j / h
        if abs(cur - prev) < 1e-5:
            break
        prev = cur
    return min(cur, max_s)


def get_projection_s(sp, tx, ty, max_s, start_s=0.0):
    """The .s of getProjection's result (the only field graphSearchDp uses)."""
    if max_s <= start_s:
        return 0.0  # State{xs(start_s), ys(start_s)}: s stays at its default
    tmp, min_s, min_dis = start_s, start_s, DBL_MAX
    while tmp <= max_s:
        x, y = _xy(sp, tmp)
        d = math.sqrt(math.pow(x - tx, 2) + math.pow(y - ty, 2))
        if d < min_dis:
            min_dis, min_s = d, tmp
        tmp += 1.0
    ex, ey = _xy(sp, max_s)
    if math.sqrt(math.pow(ex - tx, 2) + math.pow(ey - ty, 2)) < min_dis:
        return max_s
    return projection_newton(sp, tx, ty, max_s, min_s)


def _inside(dist, res, x, y):
    rows, cols = dist.shape
    return abs(x) < 0.5 * rows * res and abs(y) < 0.5 * cols * res


def _dist(dist, res, x, y):
    return float(bo.map_distance(dist, res, x, y))


def lateral_offsets():
    out, cur = [], -LATERAL_RANGE
    while cur <= LATERAL_RANGE:  # repeated addition, as the reference accumulates it (:185,213)
        out.append(cur)
        cur += LAT_SPACING
    return out


def graph_search_dp(dist, res, spline_rows, length, start_xyh):
    """Returns a dict: ok (the function's bool), and - when ok - layer_s, lower, upper (layers_s_list_ /
    layers_bounds_ after the resize), chosen (lateral index per kept layer), vehicle_l, target_s, plus the
    full tables the GPU kernel is compared against: feasible[L][J], cost[L][J], parent[L][J] (-1 = none)."""
    sp = bo.Spline2(spline_rows)
    sx, sy, sh = (float(v) for v in start_xyh)
    tmp_s = get_projection_s(sp, sx, sy, length)
    layers = []
    search_ds = LON_SPACING if length > 6 else 0.5
    while tmp_s < length:
        layers.append(tmp_s)
        tmp_s += search_ds
    layers.append(length)
    target_s = layers[-1]
    vs = layers[0]
    px, py = _xy(sp, vs)
    ph = heading(sp, vs)
    ddx, ddy = sx - px, sy - py
    vehicle_l = -ddx * math.sin(ph) + ddy * math.cos(ph)
    if abs(vehicle_l) > LATERAL_RANGE:
        return dict(ok=False, vehicle_l=vehicle_l, n_layers=len(layers))
    start_j = int((LATERAL_RANGE + vehicle_l) / LAT_SPACING)
    threshold = CAR_WIDTH / 2.0 + 0.2
    ls = lateral_offsets()
    L, J = len(layers), len(ls)
    X, Y = np.zeros((L, J)), np.zeros((L, J))
    H, DIS = np.zeros(L), np.zeros((L, J))
    feas = np.ones((L, J), dtype=bool)
    cost = np.full((L, J), DBL_MAX)
    dirs = np.zeros((L, J))
    parent = np.full((L, J), -1, dtype=np.int32)
    lo_b, up_b = np.zeros((L, J)), np.zeros((L, J))
    for i, cs in enumerate(layers):
        rx, ry = _xy(sp, cs)
        rh = heading(sp, cs)
        rk = curvature(sp, cs)
        with np.errstate(divide="ignore"):
            rr = float(np.float64(1.0) / np.float64(rk))
        H[i] = rh
        for j, cl in enumerate(ls):
            x = rx + cl * math.cos(rh + math.pi / 2)
            y = ry + cl * math.sin(rh + math.pi / 2)
            X[i, j], Y[i, j] = x, y
            d = _dist(dist, res, x, y) if _inside(dist, res, x, y) else -1.0
            DIS[i, j] = d
            if (rk < 0 and cl < rr) or (rk > 0 and cl > rr) or d < threshold:
                feas[i, j] = False
            if i == 0:
                feas[i, j] = (j == start_j)
                if j == start_j:
                    dirs[i, j], cost[i, j] = sh, 0.0
        for j in range(J):
            lo_b[i, j] = ls[j] if (j == 0 or not feas[i, j - 1] or not feas[i, j]) else lo_b[i, j - 1]
        for j in range(J - 1, -1, -1):
            up_b[i, j] = ls[j] if (j == J - 1 or not feas[i, j + 1] or not feas[i, j]) else up_b[i, j + 1]
    max_layer = 0
    for i in range(L):
        any_parent = False
        if i > 0:
            for j in range(J):
                if not feas[i, j]:
                    continue
                self_cost = 0.0
                if DIS[i, j] < SAFE_DISTANCE:
                    self_cost += (SAFE_DISTANCE - DIS[i, j]) / SAFE_DISTANCE * W_OBSTACLE
                self_cost += abs(ls[j]) / LATERAL_RANGE * W_REF_OFFSET
                best = DBL_MAX
                for p in range(J):
                    if not feas[i - 1, p] or cost[i - 1, p] == DBL_MAX:
                        continue
                    if abs(ls[p] - ls[j]) > (layers[i] - layers[i - 1]):
                        continue
                    direction = math.atan2(Y[i, j] - Y[i - 1, p], X[i, j] - X[i - 1, p])
                    edge = (abs(constrain_angle(direction - dirs[i - 1, p])) / (math.pi / 2) * W_ANGLE_CHANGE
                            + abs(constrain_angle(direction - H[i])) / (math.pi / 2) * W_REF_ANGLE_DIFF)
                    total = self_cost + edge + cost[i - 1, p]
                    if total < best:
                        best, parent[i, j], dirs[i, j] = total, p, direction
                if parent[i, j] >= 0:
                    cost[i, j] = best
                    any_parent = True
            if not any_parent:
                break
        max_layer = i
    # retrieve: first minimum of the last layer reached, then the parent chain
    jbest, cbest = -1, DBL_MAX
    for j in range(J):
        if cost[max_layer, j] < cbest:
            jbest, cbest = j, cost[max_layer, j]
    chosen, lower, upper = [], [], []
    i, j = max_layer, jbest
    while j >= 0:
        chosen.append(j)
        if i == 0:
            lower.append(-10.0)
            upper.append(10.0)
        else:
            ub, lb = CHECK_S + up_b[i, j], -CHECK_S + lo_b[i, j]
            rx, ry = _xy(sp, layers[i])
            while ub < CHECK_LIMIT:
                x, y = rx + ub * math.cos(H[i] + math.pi / 2), ry + ub * math.sin(H[i] + math.pi / 2)
                if _inside(dist, res, x, y) and _dist(dist, res, x, y) > threshold:
                    ub += CHECK_S
                else:
                    ub -= CHECK_S
                    break
            while lb > -CHECK_LIMIT:
                x, y = rx + lb * math.cos(H[i] + math.pi / 2), ry + lb * math.sin(H[i] + math.pi / 2)
                if _inside(dist, res, x, y) and _dist(dist, res, x, y) > threshold:
                    lb -= CHECK_S
                else:
                    lb += CHECK_S
                    break
            lower.append(lb)
            upper.append(ub)
        if i == 0:
            break
        j = int(parent[i, j])
        i -= 1
    chosen.reverse()
    lower.reverse()
    upper.reverse()
    n_out = len(chosen)
    return dict(ok=True, vehicle_l=vehicle_l, target_s=target_s, n_layers=L, n_out=n_out,
                layer_s=np.array(layers[:n_out]), all_layer_s=np.array(layers), lower=np.array(lower), upper=np.array(upper),
                chosen=np.array(chosen, dtype=np.int32), feasible=feas, cost=cost, parent=parent, start_j=start_j,
                max_layer=max_layer)
