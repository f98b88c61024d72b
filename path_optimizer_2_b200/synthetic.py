"""Seeded synthetic path-QP instances.

The reference ships no data, so every benchmark / parity input is generated here:
a reference line with sinusoidal curvature, knot spacing per
`ReferencePathImpl::buildReferenceFromSpline` (reference_path_impl.cpp:321-335: 0.3 m when
|k| < 0.08, 0.15 m when |k| > 0.2, linear in between), a start error, an optional
end-heading box (base_solver.cpp:250-259) and per-instance clearance bounds with
rectangular intrusions. Instance `i` of config `cfg_id` depends only on (cfg_id, i), so
any shard of a batch can be generated independently on any rank.
"""
import math

import numpy as np

from . import abi

FRONT_LENGTH = 3.9
REAR_LENGTH = -1.0


def _knot_spacing(k):
    share = min(max((abs(k) - 0.08) / 0.12, 0.0), 1.0)
    return 0.3 - share * 0.15


def make_instance(cfg_id: int, index: int, n: int, *, n_max: int = None, ragged: bool = False):
    """One instance -> (knots[9, n_max], inst[5], n_eff, ref_xyh[3, n_max])."""
    n_max = n if n_max is None else n_max
    rng = np.random.default_rng(cfg_id * 1_000_003 + index)
    a = rng.uniform(0.0, 0.15)
    f = rng.uniform(0.5, 2.0)
    phi = rng.uniform(0.0, 1.0)
    b = rng.normal(0.0, 0.03)
    n_eff = n
    if ragged:  # blocked corridors shorten n upstream (reference_path_impl.cpp:227-229)
        n_eff = int(rng.integers(max(2, n // 3), n + 1))
    length = 0.3 * n

    def kref_at(s):
        k = a * math.sin(2.0 * math.pi * (f * s / length + phi)) + b
        return min(max(k, -0.22), 0.22)

    s = np.zeros(n_max)
    kref = np.zeros(n_max)
    xyh = np.zeros((3, n_max))
    cur_s, x, y, h = 0.0, rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-math.pi, math.pi)
    for i in range(n_eff):
        k = kref_at(cur_s)
        s[i], kref[i] = cur_s, k
        xyh[:, i] = (x, y, h)
        ds = _knot_spacing(k)
        x += ds * math.cos(h + 0.5 * ds * k)
        y += ds * math.sin(h + 0.5 * ds * k)
        h += ds * k
        h = (h + math.pi) % (2 * math.pi) - math.pi
        cur_s += ds
    total = s[n_eff - 1]

    # clearance profile (left = ub > 0, right = lb < 0) with 0-2 rectangular intrusions
    wl, wr = rng.uniform(1.5, 3.5), rng.uniform(1.5, 3.5)
    pl, pr = rng.uniform(0, 2 * math.pi, size=2)
    fl, fr = rng.uniform(0.02, 0.15, size=2)
    n_intr = int(rng.integers(0, 3))
    intr = []
    for _ in range(n_intr):
        side = int(rng.integers(0, 2))
        depth = rng.uniform(0.5, 2.0)
        ln = rng.uniform(3.0, 10.0)
        start = rng.uniform(0.0, max(total - ln, 1.0))
        intr.append((side, depth, start, start + ln))

    def profile(sv):
        ub = wl + 0.8 * np.sin(fl * sv + pl)
        lb = -wr + 0.8 * np.sin(fr * sv + pr)
        for side, depth, s0, s1 in intr:
            inside = (sv >= s0) & (sv <= s1)
            if side == 0:
                ub = np.where(inside, ub - depth, ub)
            else:
                lb = np.where(inside, lb + depth, lb)
        # always leave >= 0.2 m
        mid = 0.5 * (ub + lb)
        ub = np.maximum(ub, mid + 0.1)
        lb = np.minimum(lb, mid - 0.1)
        return lb, ub

    sv = s[:n_eff]
    f_lb, f_ub = profile(sv + FRONT_LENGTH)
    r_lb, r_ub = profile(sv + REAR_LENGTH)

    knots = np.zeros((abi.NFIELDS, n_max))
    knots[abi.F_S] = s
    knots[abi.F_KREF] = kref
    knots[abi.F_L] = 0.0      # first solve linearises about the reference line
    knots[abi.F_PSI] = 0.0    # (path_optimizer.cpp:128-137)
    knots[abi.F_K] = kref
    knots[abi.F_B0_LB, :n_eff] = f_lb
    knots[abi.F_B0_UB, :n_eff] = f_ub
    knots[abi.F_B1_LB, :n_eff] = r_lb
    knots[abi.F_B1_UB, :n_eff] = r_ub

    inst = np.zeros(abi.NINST)
    inst[abi.I_L0] = rng.uniform(-0.5, 0.5)
    inst[abi.I_PSI0] = rng.uniform(-0.1, 0.1)
    inst[abi.I_K0] = kref[0]
    if rng.uniform() < 0.5:
        inst[abi.I_EPSI_LO], inst[abi.I_EPSI_HI] = -abi.INFTY, abi.INFTY
    else:
        e = rng.uniform(-0.1, 0.1)
        inst[abi.I_EPSI_LO], inst[abi.I_EPSI_HI] = e - 0.087, e + 0.087
    return knots, inst, n_eff, xyh


def make_batch(cfg_id: int, batch: int, n: int, *, first: int = 0, n_max: int = None,
               ragged: bool = False, with_ref: bool = False):
    """Instances [first, first+batch) of config `cfg_id` as an abi.HostBatch."""
    n_max = n if n_max is None else n_max
    knots = np.zeros((batch, abi.NFIELDS, n_max))
    inst = np.zeros((batch, abi.NINST))
    ns = np.zeros(batch, dtype=np.int32)
    ref = np.zeros((batch, 3, n_max)) if with_ref else None
    for b in range(batch):
        k, i, ne, xyh = make_instance(cfg_id, first + b, n, n_max=n_max, ragged=ragged)
        knots[b], inst[b], ns[b] = k, i, ne
        if with_ref:
            ref[b] = xyh
    hb = abi.HostBatch(knots, inst, ns)
    return (hb, ref) if with_ref else hb


# BASELINE.json configs that are synthetic (configs[2..4]); configs[0..1] need the map
# fixture and are built by tests/fixtures (shared-map bounds).
CONFIGS = {
    3: dict(batch=8192, n=240),
    4: dict(batch=65536, n=240),
    5: dict(batch=4096, n=240, max_iter=50, ticks=20),
}


def shift_window(ext_knots, inst, sol, tick, n):
    """Receding-horizon bookkeeping of BASELINE configs[4]: the planning
    window advances by one knot per tick along an extended reference `ext_knots[B, 9, n+T]`
    (only its S / KREF / bounds rows are used); the new linearisation point is the previous
    solution shifted by one knot (last knot repeated) and the new x0 is the previous solution at
    knot 1. Works on numpy arrays and on torch tensors (same slicing)."""
    lo = tick
    knots = ext_knots[:, :, lo:lo + n].clone() if hasattr(ext_knots, "clone") else ext_knots[:, :, lo:lo + n].copy()
    for f_dst, f_src in ((abi.F_L, 0), (abi.F_PSI, 1), (abi.F_K, 2)):
        knots[:, f_dst, :n - 1] = sol[:, f_src, 1:n]
        knots[:, f_dst, n - 1] = sol[:, f_src, n - 1]
    new_inst = inst.clone() if hasattr(inst, "clone") else inst.copy()
    new_inst[:, abi.I_L0] = sol[:, 0, 1]
    new_inst[:, abi.I_PSI0] = sol[:, 1, 1]
    new_inst[:, abi.I_K0] = sol[:, 2, 1]
    return knots, new_inst
