"""Smoother QPs: the CUDA library's routine on the CPU (tests/emu/smoother_driver.cpp)
and on the GPU (`-m gpu`) against the generic OSQP restatement (oracle/smoother_oracle.py, osqp_generic.py).

Parity bar (FP64 on both sides, same algorithm, different linear solver - reduced banded LDL' vs SuperLU on
the KKT matrix): identical status, iteration count and number of rho updates; the solution within 1e-6 of the
oracle's and passing OSQP's termination test (eps 1e-3, OSQP's default: the reference overrides nothing here)
evaluated on the oracle-assembled QP."""
import os
import re

import numpy as np
import pytest

from oracle import osqp_generic, smoother_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw_reference(seed, p, ds=1.0):
    """A wiggly raw reference resampled at `ds` (what segmentRawReference hands osqpSmooth): heading = integrated
    curvature + noise, curvature list = its finite difference."""
    rng = np.random.default_rng(seed)
    s = ds * np.arange(p)
    kap = 0.08 * np.sin(2 * np.pi * (rng.uniform(0.5, 2.0) * s / s[-1] + rng.uniform())) + rng.normal(0, 0.03, p)
    a = np.concatenate(([rng.uniform(-1, 1)], np.cumsum(kap[:-1] * ds))) + rng.normal(0, 0.02, p)
    x = np.concatenate(([rng.uniform(-5, 5)], np.cumsum(ds * np.cos(a[:-1])))) + rng.normal(0, 0.05, p)
    y = np.concatenate(([rng.uniform(-5, 5)], np.cumsum(ds * np.sin(a[:-1])))) + rng.normal(0, 0.05, p)
    return x, y, a, kap, s


def corridor(seed, p):
    """DP-style layers 1.5 m apart with a corridor whose bounds become ACTIVE (they exclude l = 0 in places)."""
    rng = np.random.default_rng(seed)
    layer_s = 1.5 * np.arange(p) + rng.uniform(0, 1)
    centre = 1.5 * np.sin(2 * np.pi * rng.uniform(0.5, 1.5) * np.arange(p) / p + rng.uniform(0, 6))
    half = rng.uniform(0.3, 1.2, p)
    lower, upper = centre - half, centre + half
    lower[0], upper[0] = -10.0, 10.0
    return layer_s, lower, upper, float(rng.uniform(-0.8, 0.8))


def _check_tension(res, x, y, a, k, s):
    ok, rx, ry, rs, g = so.osqp_smooth(x, y, a, k, s)
    xs, _, _ = g.solution()
    assert res["status"] == g.status and res["iters"] == g.iters, (res["status"], res["iters"], g.status, g.iters)
    assert np.max(np.abs(res["x_full"] - xs)) < 1e-6
    assert np.allclose(res["x"], rx, atol=1e-6) and np.allclose(res["y"], ry, atol=1e-6) and np.allclose(res["s"], rs, atol=1e-6)
    H, q, A, lo, up = so.assemble(x, y, a, k, s)
    Ax = A @ res["x_full"]
    assert np.max(np.abs(Ax - lo)) < 1e-3 + 1e-3 * max(np.max(np.abs(Ax)), np.max(np.abs(lo)))  # primal part of the termination test
    return g


def _check_post(res, layer_s, lower, upper, vl):
    ok, off, g = so.post_smooth(layer_s, lower, upper, vl)
    xs, _, _ = g.solution()
    assert res["status"] == g.status and res["iters"] == g.iters, (res["status"], res["iters"], g.status, g.iters)
    assert np.max(np.abs(res["x_full"] - xs)) < 1e-6 and np.allclose(res["offsets"], off, atol=1e-6)
    return g


def test_post_oracle_closed_forms():
    p = 20
    layer_s = 1.5 * np.arange(p)
    ok, off, g = so.post_smooth(layer_s, np.full(p, -2.0), np.full(p, 2.0), 0.0)
    assert ok and np.max(np.abs(off)) < 1e-3                      # on the reference already: stays there
    H, q, A, lo, up = so.assemble_post(layer_s, np.full(p, -2.0), np.full(p, 2.0), 0.5)
    assert H.shape == (3 * p, 3 * p) and A.shape == (3 * p - 2, 3 * p) and A.nnz == p + 6 * (p - 1)
    assert lo[0] == up[0] == 0.5 and np.all(lo[p:] == 0) and np.all(up[p:] == 0)
    ok, off, g = so.post_smooth(layer_s, np.full(p, 0.3), np.full(p, 2.0), 1.0)   # pushed to one side
    assert ok and abs(off[0] - 1.0) < 2e-3 and np.all(off[1:] > 0.3 - 2e-3)


@pytest.mark.parametrize("seed,p", [(1, 12), (2, 40), (3, 75), (4, 120)])
def test_kernel_source_on_cpu_tension(seed, p):
    from tests.emu import smoother_emu as se
    x, y, a, k, s = raw_reference(seed, p)
    g = _check_tension(se.tension(x, y, a, k, s), x, y, a, k, s)
    assert g.status == osqp_generic.SOLVED


@pytest.mark.parametrize("seed,p", [(1, 8), (2, 33), (3, 60), (4, 100)])
def test_kernel_source_on_cpu_post(seed, p):
    from tests.emu import smoother_emu as se
    layer_s, lower, upper, vl = corridor(seed, p)
    g = _check_post(se.post(layer_s, lower, upper, vl), layer_s, lower, upper, vl)
    assert g.status == osqp_generic.SOLVED and g.iters >= 50   # active bounds: more than one check interval


def test_smoother_symbols_declared_and_exported():
    from path_optimizer_2_b200 import smoother
    L = smoother._lib()
    with open(os.path.join(ROOT, "include", "pqp_smoother.h")) as f:
        declared = set(m.group(1) for m in re.finditer(r"^\s*(?:int|void|const char \*)\s*(pqp_\w+)\s*\(", f.read(), re.M))
    assert declared == set(smoother.EXPORTED_SYMBOLS) and all(hasattr(L, s) for s in declared)


@pytest.mark.gpu
def test_gpu_smoother_matches_oracle():
    from path_optimizer_2_b200 import smoother
    sm = smoother.Smoother(p_max=128, batch_max=64)
    cases = [raw_reference(100 + b, 20 + 3 * b) for b in range(24)]
    r = sm.tension(*[[c[i] for c in cases] for i in range(5)], full=True)
    for b, (x, y, a, k, s) in enumerate(cases):
        p = len(x)
        res = dict(status=r["status"][b], iters=r["iters"][b], x=r["x"][b, :p], y=r["y"][b, :p], s=r["s"][b, :p],
                   x_full=r["x_full"][b, :4 * p - 1])
        _check_tension(res, x, y, a, k, s)
    t_ms = sm.last_kernel_ms
    cors = [corridor(200 + b, 10 + 4 * b) for b in range(24)]
    r = sm.post([c[0] for c in cors], [c[1] for c in cors], [c[2] for c in cors], [c[3] for c in cors], full=True)
    for b, (layer_s, lower, upper, vl) in enumerate(cors):
        p = len(layer_s)
        xf = np.concatenate([r["x_full"][b, j * p:(j + 1) * p] for j in range(3)])
        _check_post(dict(status=r["status"][b], iters=r["iters"][b], offsets=r["offsets"][b, :p], x_full=xf), layer_s, lower, upper, vl)
    print("smoother kernels: tension %.3f ms, post %.3f ms for 24 QPs each" % (t_ms, sm.last_kernel_ms))
    sm.close()
