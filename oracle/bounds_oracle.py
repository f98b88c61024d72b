"""bounds_oracle.py — CPU restatement (numpy, FP64) of the reference's clearance-bounds front end
(SURVEY.md §8 row f-1). TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's checker / cpu_baseline legs; the product path is the CUDA kernel in
path_optimizer_2_b200/csrc/pqp_bounds.cu and never calls this.

PARITY UNPINNED: the reference has no test or golden vector for this path and cannot be built
here (grid_map, OpenCV-through-ROS, glog are absent), so this restatement is checked only against
closed-form cases (tests/test_bounds.py: analytic corridors, circles) — not against reference output.

Follows, relative to /root/reference/:
  map_distance            src/tools/Map.cpp:16-22 (grid_map INTER_LINEAR lookup; grid_map itself is EXT)
  natural_spline          src/tools/spline.cpp:163-247 (tk::spline::set_points, default boundary)
  spline_eval             src/tools/spline.cpp:252-330 (operator(), deriv)
  directional_projection  src/tools/tools.cpp:156-189
  clearance               src/data_struct/reference_path_impl.cpp:232-312
  update_bounds           src/data_struct/reference_path_impl.cpp:177-230
  build_states            src/data_struct/reference_path_impl.cpp:314-338, src/tools/tools.cpp:32-44
"""
import math

import numpy as np

FRONT_LENGTH, REAR_LENGTH, CAR_WIDTH, SAFETY_MARGIN, EPSILON = 3.9, -1.0, 2.0, 0.3, 1e-6


def map_distance(dist, res, x, y, cx=0.0, cy=0.0):
    """`dist`: float32 [rows, cols] distance layer; cell (i, j) centred at
    (cx + Lx/2 - (i + 1/2) res, cy + Ly/2 - (j + 1/2) res). 0 outside the map (Map.cpp:19-21)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    rows, cols = dist.shape
    hx, hy = 0.5 * rows * res, 0.5 * cols * res
    dx, dy = x - cx, y - cy
    with np.errstate(invalid="ignore"):
        inside = (np.abs(dx) < hx) & (np.abs(dy) < hy)
    dxs, dys = np.where(inside, dx, 0.0), np.where(inside, dy, 0.0)
    fi = (hx - dxs) / res - 0.5
    fj = (hy - dys) / res - 0.5
    i0 = np.clip(np.floor(fi).astype(np.int64), 0, rows - 2)
    j0 = np.clip(np.floor(fj).astype(np.int64), 0, cols - 2)
    ti = np.clip(fi - i0, 0.0, 1.0)
    tj = np.clip(fj - j0, 0.0, 1.0)
    d00 = dist[i0, j0].astype(np.float64)
    d01 = dist[i0, j0 + 1].astype(np.float64)
    d10 = dist[i0 + 1, j0].astype(np.float64)
    d11 = dist[i0 + 1, j0 + 1].astype(np.float64)
    v = d00 * (1.0 - ti) * (1.0 - tj) + d10 * ti * (1.0 - tj) + d01 * (1.0 - ti) * tj + d11 * ti * tj
    return np.where(inside, v, 0.0)


def natural_spline(sx, y):
    """tk::spline::set_points with the default boundary (second derivative 0 at both ends, no forced
    linear extrapolation): returns (a, b, c) with the reference's right-end entries
    (a[k-1] = 0, c[k-1] = f'(x_{k-1}), spline.cpp:241-246)."""
    sx = np.asarray(sx, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    k = len(sx)
    assert k > 2 and np.all(np.diff(sx) > 0)
    h = np.diff(sx)
    lo, di, up, rhs = np.zeros(k), np.zeros(k), np.zeros(k), np.zeros(k)
    lo[1:k - 1] = h[:-1] / 3.0
    di[1:k - 1] = 2.0 / 3.0 * (sx[2:] - sx[:-2])
    up[1:k - 1] = h[1:] / 3.0
    rhs[1:k - 1] = (y[2:] - y[1:-1]) / h[1:] - (y[1:-1] - y[:-2]) / h[:-1]
    di[0] = di[k - 1] = 2.0  # 2 b = f'' = 0
    # Thomas algorithm (the reference LU-factors the same tridiagonal band matrix)
    cp, dp = np.zeros(k), np.zeros(k)
    cp[0], dp[0] = up[0] / di[0], rhs[0] / di[0]
    for i in range(1, k):
        den = di[i] - lo[i] * cp[i - 1]
        cp[i] = up[i] / den
        dp[i] = (rhs[i] - lo[i] * dp[i - 1]) / den
    b = np.zeros(k)
    b[k - 1] = dp[k - 1]
    for i in range(k - 2, -1, -1):
        b[i] = dp[i] - cp[i] * b[i + 1]
    a, c = np.zeros(k), np.zeros(k)
    a[:-1] = (b[1:] - b[:-1]) / (3.0 * h)
    c[:-1] = (y[1:] - y[:-1]) / h - (2.0 * b[:-1] + b[1:]) * h / 3.0
    hl = h[-1]
    c[k - 1] = 3.0 * a[k - 2] * hl * hl + 2.0 * b[k - 2] * hl + c[k - 2]
    return a, b, c


def spline_eval(sx, a, b, c, y, s):
    """(value, first, second derivative) at s (vectorised), with the reference's segment rule
    (lower_bound - 1, clamped at 0) and its extrapolation formulas."""
    s = np.asarray(s, dtype=np.float64)
    k = len(sx)
    idx = np.maximum(np.searchsorted(sx, s, side="left") - 1, 0)
    idx = np.minimum(idx, k - 1)
    h = s - sx[idx]
    v = ((a[idx] * h + b[idx]) * h + c[idx]) * h + y[idx]
    d1 = (3.0 * a[idx] * h + 2.0 * b[idx]) * h + c[idx]
    d2 = 6.0 * a[idx] * h + 2.0 * b[idx]
    with np.errstate(invalid="ignore"):
        left, right = s < sx[0], s > sx[k - 1]
    hl, hr = s - sx[0], s - sx[k - 1]
    v = np.where(left, (b[0] * hl + c[0]) * hl + y[0], v)
    d1 = np.where(left, 2.0 * b[0] * hl + c[0], d1)
    d2 = np.where(left, 2.0 * b[0] * hl, d2)
    v = np.where(right, (b[k - 1] * hr + c[k - 1]) * hr + y[k - 1], v)
    d1 = np.where(right, 2.0 * b[k - 1] * hr + c[k - 1], d1)
    d2 = np.where(right, 2.0 * b[k - 1] + 0.0 * hr, d2)
    return v, d1, d2


class Spline2:
    """x(s), y(s) over common abscissae; rows as include/pqp_bounds.h packs them."""

    def __init__(self, rows):
        self.sx, self.xa, self.xb, self.xc, self.xy, self.ya, self.yb, self.yc, self.yy = (np.asarray(r, dtype=np.float64) for r in rows)

    def x(self, s):
        return spline_eval(self.sx, self.xa, self.xb, self.xc, self.xy, s)

    def y(self, s):
        return spline_eval(self.sx, self.ya, self.yb, self.yc, self.yy, s)


def directional_projection(sp, tx, ty, angle, max_s, hint_s):
    """tools.cpp:156-189, vectorised; returns the foot point (x, y)."""
    with np.errstate(all="ignore"):
        cur = np.where(max_s < hint_s, max_s, hint_s).astype(np.float64)
        prev = cur.copy()
        v1, v2 = np.sin(angle), -np.cos(angle)
        active = np.ones(cur.shape, dtype=bool)
        for _ in range(20):
            x, dx, ddx = sp.x(cur)
            y, dy, ddy = sp.y(cur)
            p1 = v1 * (x - tx) + v2 * (y - ty)
            p2 = v1 * dx + v2 * dy
            j = p1 * p2
            h = p1 * (v1 * ddx + v2 * ddy) + p2 * p2
            cur = np.where(active, cur - j / h, cur)
            done = np.abs(cur - prev) < 1e-5
            prev = np.where(active, cur, prev)
            active &= ~done
        cur = np.where(max_s < cur, max_s, cur)
        return sp.x(cur)[0], sp.y(cur)[0]


def _constrain(a):
    a = np.array(a, dtype=np.float64)
    for _ in range(8):
        a = np.where(a > math.pi, a - 2 * math.pi, a)
    for _ in range(8):
        a = np.where(a < -math.pi, a + 2 * math.pi, a)
    return a


def clearance(dist, res, x, y, heading, car_width=CAR_WIDTH, safety_margin=SAFETY_MARGIN, lookup=None):
    """reference_path_impl.cpp:232-312, vectorised over states; returns (lb, ub) = (right, left)."""
    look = lookup if lookup is not None else (lambda px, py: map_distance(dist, res, px, py))
    x, y, heading = (np.asarray(v, dtype=np.float64) for v in (x, y, heading))
    delta_s, search_radius, smaller_ds, min_space = 0.3, 0.5, 0.05, 0.2
    la, ra = _constrain(heading + math.pi / 2), _constrain(heading - math.pi / 2)
    cl, sl, cr, sr = np.cos(la), np.sin(la), np.cos(ra), np.sin(ra)
    n = int(6.0 / delta_s)
    ok = look(x, y) > search_radius

    def march(c, s_):
        dist_s = np.zeros_like(x)
        active = np.ones(x.shape, dtype=bool)
        for _ in range(n):
            dist_s = np.where(active, dist_s + delta_s, dist_s)
            hit = look(x + dist_s * c, y + dist_s * s_) < search_radius
            active &= ~hit
        return dist_s

    right_s, left_s = march(cr, sr), march(cl, sl)
    right_b, left_b = -(right_s - delta_s), left_s - delta_s
    fine = int(delta_s / smaller_ds)
    active = np.ones(x.shape, dtype=bool)
    for _ in range(1, fine):
        cand = left_b + smaller_ds
        hit = look(x + cand * cl, y + cand * sl) < search_radius
        left_b = np.where(active, np.where(hit, cand - smaller_ds, cand), left_b)
        active &= ~hit
    active = np.ones(x.shape, dtype=bool)
    for _ in range(1, fine):
        cand = right_b - smaller_ds
        hit = look(x + cand * cr, y + cand * sr) < search_radius  # negative bound x right-hand direction (:288-291)
        right_b = np.where(active, np.where(hit, cand + smaller_ds, cand), right_b)
        active &= ~hit
    diff = car_width * 0.5 - search_radius
    left_b, right_b = left_b - diff, right_b + diff
    blocked = left_b < right_b
    space = left_b - right_b
    margin = np.minimum(safety_margin, np.maximum(0.0, (space - min_space) / 2.0))
    bad = ~ok | blocked
    return np.where(bad, 0.0, right_b + margin), np.where(bad, 0.0, left_b - margin)


def update_bounds(dist, res, spline_rows, s, x, y, heading, *, front_length=FRONT_LENGTH, rear_length=REAR_LENGTH,
                  car_width=CAR_WIDTH, safety_margin=SAFETY_MARGIN, epsilon=EPSILON, lookup=None):
    """updateBoundsImproved for one path: returns (bounds[6, n], n_valid) with rows front lb/ub,
    rear lb/ub, centre lb/ub; n_valid = index of the first blocked state (n when none)."""
    s, x, y, heading = (np.asarray(v, dtype=np.float64) for v in (s, x, y, heading))
    sp = Spline2(spline_rows)
    out = np.zeros((6, len(s)))
    ch, sh = np.cos(heading), np.sin(heading)
    for row, length in ((0, front_length), (2, rear_length)):
        ax, ay = x + length * ch, y + length * sh
        px, py = directional_projection(sp, ax, ay, heading + math.pi / 2, s + 5.0, s + length)
        lb, ub = clearance(dist, res, px, py, heading, car_width, safety_margin, lookup)
        with np.errstate(invalid="ignore"):
            offset = -(px - ax) * sh + (py - ay) * ch
        out[row], out[row + 1] = lb + offset, ub + offset
    out[4], out[5] = clearance(dist, res, x, y, heading, car_width, safety_margin, lookup)
    with np.errstate(invalid="ignore"):
        blocked = (np.abs(out[1] - out[0]) < epsilon) | (np.abs(out[3] - out[2]) < epsilon)
    hits = np.nonzero(blocked)[0]
    return out, int(hits[0]) if len(hits) else len(s)


def build_states(spline_rows, max_s, ds_small=0.15, ds_large=0.3, dynamic=True):
    """buildReferenceFromSpline for one path: (s, x, y, heading, curvature) arrays, uncapped."""
    sp = Spline2(spline_rows)
    out = []
    tmp_s = 0.0
    while tmp_s <= max_s:
        x, dx, ddx = (float(v) for v in sp.x(tmp_s))
        y, dy, ddy = (float(v) for v in sp.y(tmp_s))
        h = math.atan2(dy, dx)
        k = (dx * ddy - dy * ddx) / math.pow(math.pow(dx, 2) + math.pow(dy, 2), 1.5)
        out.append((tmp_s, x, y, h, k))
        if dynamic:
            ak = abs(k)
            share = 1.0 if ak > 0.2 else (0.0 if ak < 0.08 else (ak - 0.08) / (0.2 - 0.08))
            tmp_s += ds_large - share * (ds_large - ds_small)
        else:
            tmp_s += ds_large
    return tuple(np.array(c) for c in zip(*out))
