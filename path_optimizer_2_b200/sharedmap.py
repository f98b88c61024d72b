"""Shared-obstacle-map workload (BASELINE configs[0] and configs[1]).

Input generation only (host side, numpy): the solver consumes per-knot clearance bounds; in the
reference those come from ray-marching a distance field of `gridmap.png`. This module restates
that front end on the CPU so that the shared-map configurations have realistic inputs:

  distance field      src/test/demo.cpp:98-113  (cv2.distanceTransform(DIST_L2, MASK_PRECISE) * 0.2 m)
  map lookup          src/tools/Map.cpp:16-22   (bilinear inside the map, 0 outside)  [grid_map: EXT]
  reference spline    src/tools/spline.cpp:163-254  (natural cubic spline over arc length)
  knots               src/data_struct/reference_path_impl.cpp:314-338, src/tools/tools.cpp:32-44
  anchors + bounds    reference_path_impl.cpp:177-230 (front/rear anchors projected onto the spline
                      along the knot normal, tools.cpp:156-189) and :232-312 (clearance ray-march)

The GPU version of this front end is SURVEY.md §8f-1 ("next"); nothing here is on the solve path.
"""
import math
import os

import numpy as np

from . import abi

RES = 0.2
FRONT_LENGTH, REAR_LENGTH = 3.9, -1.0
CAR_WIDTH, SAFETY_MARGIN = 2.0, 0.3
_MAP_PNG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gridmap.png")


class DistanceMap:
    """grid_map "distance" layer of the demo: cell (i, j) = image pixel (row i, col j), its
    centre at (+Lx/2 - (i + 1/2) res, +Ly/2 - (j + 1/2) res) with Lx = rows * res, Ly = cols * res."""

    def __init__(self, png_path=_MAP_PNG):
        if not os.path.exists(png_path):
            raise FileNotFoundError(png_path)
        try:
            import cv2
            img = cv2.imread(png_path, cv2.IMREAD_GRAYSCALE)
            free = (img > 127).astype(np.uint8)  # OCCUPY = 0, FREE = 255 (demo.cpp:103-106)
            dist = cv2.distanceTransform(free, cv2.DIST_L2, cv2.DIST_MASK_PRECISE)
        except ImportError:  # same exact Euclidean transform without OpenCV
            from PIL import Image
            from scipy.ndimage import distance_transform_edt
            img = np.asarray(Image.open(png_path).convert("L"))
            dist = distance_transform_edt(img > 127)
        self.dist = dist.astype(np.float64) * RES
        self.rows, self.cols = self.dist.shape
        self.lx, self.ly = self.rows * RES, self.cols * RES

    def lookup(self, x, y):
        """Map::getObstacleDistance, vectorised: bilinear between the four nearest cell centres."""
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        fi = (self.lx / 2 - x) / RES - 0.5  # fractional row index
        fj = (self.ly / 2 - y) / RES - 0.5
        inside = (np.abs(x) < self.lx / 2) & (np.abs(y) < self.ly / 2)
        i0 = np.clip(np.floor(fi).astype(np.int64), 0, self.rows - 2)
        j0 = np.clip(np.floor(fj).astype(np.int64), 0, self.cols - 2)
        ti = np.clip(fi - i0, 0.0, 1.0)
        tj = np.clip(fj - j0, 0.0, 1.0)
        d = self.dist
        v = (d[i0, j0] * (1 - ti) * (1 - tj) + d[i0 + 1, j0] * ti * (1 - tj) +
             d[i0, j0 + 1] * (1 - ti) * tj + d[i0 + 1, j0 + 1] * ti * tj)
        return np.where(inside, v, 0.0)


def _constrain(a):
    return (a + math.pi) % (2 * math.pi) - math.pi


def clearance(dmap, x, y, heading):
    """ReferencePathImpl::getClearanceWithDirectionStrict (reference_path_impl.cpp:232-312),
    vectorised over points. Returns (left = ub, right = lb)."""
    x, y, heading = (np.asarray(v, dtype=np.float64) for v in (x, y, heading))
    delta_s, search_radius = 0.3, 0.5
    nstep = int(6.0 / delta_s)
    la, ra = _constrain(heading + math.pi / 2), _constrain(heading - math.pi / 2)
    ok = dmap.lookup(x, y) > search_radius

    def march(angle):
        s = np.zeros_like(x)
        active = np.ones(x.shape, dtype=bool)
        for _ in range(nstep):
            s = np.where(active, s + delta_s, s)
            c = dmap.lookup(x + s * np.cos(angle), y + s * np.sin(angle))
            active &= ~(c < search_radius)
        return s

    right_s, left_s = march(ra), march(la)
    left, right = left_s - delta_s, -(right_s - delta_s)
    smaller = 0.05
    act = np.ones(x.shape, dtype=bool)
    for _ in range(1, int(delta_s / smaller)):  # refine forward (:272-283)
        cand = np.where(act, left + smaller, left)
        hit = dmap.lookup(x + cand * np.cos(la), y + cand * np.sin(la)) < search_radius
        left = np.where(act & ~hit, cand, left)
        act &= ~hit
    act = np.ones(x.shape, dtype=bool)
    for _ in range(1, int(delta_s / smaller)):  # (:284-295; negative bound x right-hand direction, as there)
        cand = np.where(act, right - smaller, right)
        hit = dmap.lookup(x + cand * np.cos(ra), y + cand * np.sin(ra)) < search_radius
        right = np.where(act & ~hit, cand, right)
        act &= ~hit
    diff = CAR_WIDTH * 0.5 - search_radius
    left, right = left - diff, right + diff
    blocked = left < right
    space = left - right
    margin = np.minimum(SAFETY_MARGIN, np.maximum(0.0, (space - 0.2) / 2.0))
    left, right = left - margin, right + margin
    bad = ~ok | blocked
    return np.where(bad, 0.0, left), np.where(bad, 0.0, right)


class SplinePath:
    """x(s), y(s) natural cubic splines (tk::spline defaults) + tools.cpp helpers."""

    def __init__(self, s, x, y):
        from scipy.interpolate import CubicSpline
        self.xs, self.ys = CubicSpline(s, x, bc_type="natural"), CubicSpline(s, y, bc_type="natural")
        self.max_s = float(s[-1])

    def heading(self, s):  # tools.cpp:32-36
        return np.arctan2(self.ys(s, 1), self.xs(s, 1))

    def curvature(self, s):  # tools.cpp:38-44
        dx, dy, ddx, ddy = self.xs(s, 1), self.ys(s, 1), self.xs(s, 2), self.ys(s, 2)
        return (dx * ddy - dy * ddx) / np.power(dx * dx + dy * dy, 1.5)

    def directional_projection(self, tx, ty, angle, hint_s):  # tools.cpp:156-189, vectorised
        cur = np.minimum(np.asarray(hint_s, dtype=np.float64), self.max_s)
        prev = cur.copy()
        v1, v2 = np.sin(angle), -np.cos(angle)
        act = np.ones(cur.shape, dtype=bool)
        for _ in range(20):
            xv, yv = self.xs(cur), self.ys(cur)
            dx, dy, ddx, ddy = self.xs(cur, 1), self.ys(cur, 1), self.xs(cur, 2), self.ys(cur, 2)
            p1 = v1 * (xv - tx) + v2 * (yv - ty)
            p2 = v1 * dx + v2 * dy
            h = p1 * (v1 * ddx + v2 * ddy) + p2 * p2
            step = np.where(np.abs(h) > 1e-12, p1 * p2 / np.where(h == 0, 1.0, h), 0.0)
            cur = np.where(act, cur - step, cur)
            act &= ~(np.abs(cur - prev) < 1e-5)
            prev = cur.copy()
        cur = np.minimum(cur, self.max_s)
        return self.xs(cur), self.ys(cur)


def build_knots(path, n):
    """buildReferenceFromSpline(0.15, 0.3): first n knots, or None if the spline is too short."""
    s_list, s = [], 0.0
    while s <= path.max_s and len(s_list) < n:
        s_list.append(s)
        k = abs(float(path.curvature(s)))
        share = 1.0 if k > 0.2 else (0.0 if k < 0.08 else (k - 0.08) / 0.12)
        s += 0.3 - share * 0.15
    if len(s_list) < n:
        return None
    s = np.array(s_list)
    return s, path.xs(s), path.ys(s), path.heading(s), path.curvature(s)


def bounds_for(dmap, path, s, x, y, h):
    """updateBoundsImproved (reference_path_impl.cpp:177-230) for all knots; returns
    (f_lb, f_ub, r_lb, r_ub, c_lb, c_ub, blocked)."""
    out = {}
    for name, length in (("f", FRONT_LENGTH), ("r", REAR_LENGTH)):
        ax, ay = x + length * np.cos(h), y + length * np.sin(h)
        px, py = path.directional_projection(ax, ay, h + math.pi / 2, s + length)
        ub, lb = clearance(dmap, px, py, h)
        # offset of the projected anchor in the raw anchor's frame (global2Local(...).y)
        off = -(px - ax) * np.sin(h) + (py - ay) * np.cos(h)
        out[name] = (lb + off, ub + off)
    cub, clb = clearance(dmap, x, y, h)
    blocked = (np.abs(out["f"][1] - out["f"][0]) < 1e-6) | (np.abs(out["r"][1] - out["r"][0]) < 1e-6)
    return out["f"][0], out["f"][1], out["r"][0], out["r"][1], clb, cub, blocked


def _walk_free_line(dmap, rng, x0, y0, total, step=1.5, min_clear=1.6):
    """Stand-in for the reference's hybrid-A* front end (out of scope): walk from
    (x0, y0) in 1.5 m arcs, each time taking the curvature whose 6 m look-ahead keeps the most
    clearance, and return control points (s, x, y) or None when the corridor closes."""
    cand = np.linspace(-0.12, 0.12, 9)
    look = step * np.arange(1, 5)
    best_h, best_c = 0.0, -1.0
    for h in rng.uniform(-math.pi, math.pi, size=8):  # start along the most open direction
        c = float(dmap.lookup(x0 + look * math.cos(h), y0 + look * math.sin(h)).min())
        if c > best_c:
            best_h, best_c = h, c
    x, y, h, k_prev = x0, y0, best_h, 0.0
    xs, ys = [x], [y]
    for _ in range(int(math.ceil(total / step))):
        hh = h + 0.5 * cand[:, None] * look[None, :]
        px, py = x + look[None, :] * np.cos(hh), y + look[None, :] * np.sin(hh)
        score = np.minimum(dmap.lookup(px, py).min(axis=1), 2.5) - 4.0 * np.abs(cand - k_prev) - 2.0 * np.abs(cand)
        k = float(cand[int(np.argmax(score))])
        x, y = x + step * math.cos(h + 0.5 * step * k), y + step * math.sin(h + 0.5 * step * k)
        h, k_prev = h + step * k, k
        if dmap.lookup(x, y) < min_clear:
            return None
        xs.append(x)
        ys.append(y)
    xs, ys = np.array(xs), np.array(ys)
    sc = np.concatenate(([0.0], np.cumsum(np.hypot(np.diff(xs), np.diff(ys)))))
    return sc, xs, ys


def make_instance(dmap, seed, n):
    """One unblocked n-knot instance on the shared map (rejection sampling), as
    (knots[9, n], inst[5], ref_xyh[3, n])."""
    rng = np.random.default_rng(seed)
    for _ in range(400):
        x0 = rng.uniform(-dmap.lx / 2 + 5, dmap.lx / 2 - 5)
        y0 = rng.uniform(-dmap.ly / 2 + 5, dmap.ly / 2 - 5)
        if dmap.lookup(x0, y0) < 2.0:
            continue
        line = _walk_free_line(dmap, rng, x0, y0, 0.3 * n + 8.0)
        if line is None:
            continue
        sc, xx, yy = line
        path = SplinePath(sc, xx, yy)
        kn = build_knots(path, n)
        if kn is None:
            continue
        s, x, y, h, k = kn
        flb, fub, rlb, rub, clb, cub, blocked = bounds_for(dmap, path, s, x, y, h)
        if np.any(blocked) or np.any(fub - flb < 0.05) or np.any(rub - rlb < 0.05):
            continue
        knots = np.zeros((abi.NFIELDS, n))
        knots[abi.F_S], knots[abi.F_KREF], knots[abi.F_K] = s, k, k
        knots[abi.F_B0_LB], knots[abi.F_B0_UB] = flb, fub
        knots[abi.F_B1_LB], knots[abi.F_B1_UB] = rlb, rub
        inst = np.zeros(abi.NINST)
        inst[abi.I_L0] = rng.uniform(-0.3, 0.3)
        inst[abi.I_PSI0] = rng.uniform(-0.08, 0.08)
        inst[abi.I_K0] = k[0]
        inst[abi.I_EPSI_LO], inst[abi.I_EPSI_HI] = -abi.INFTY, abi.INFTY
        if rng.uniform() < 0.5:
            e = rng.uniform(-0.05, 0.05)
            inst[abi.I_EPSI_LO], inst[abi.I_EPSI_HI] = e - 0.087, e + 0.087
        return knots, inst, np.stack((x, y, h))
    raise RuntimeError("no free reference line found for seed %d" % seed)


def make_batch(batch, n=120, first=0, cfg_id=2, with_ref=False, dmap=None):
    """BASELINE configs[1]: `batch` paths of n knots through the shared obstacle map."""
    dmap = dmap or DistanceMap()
    knots = np.zeros((batch, abi.NFIELDS, n))
    inst = np.zeros((batch, abi.NINST))
    ref = np.zeros((batch, 3, n))
    for b in range(batch):
        knots[b], inst[b], ref[b] = make_instance(dmap, cfg_id * 1_000_003 + first + b, n)
    hb = abi.HostBatch(knots, inst, np.full(batch, n, dtype=np.int32))
    return (hb, ref) if with_ref else hb
