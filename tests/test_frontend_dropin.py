"""The C++ drop-in for ReferencePathImpl::buildReferenceFromSpline / updateBoundsImproved
(include/pqp_reference_path.hpp), driven like the reference drives those methods.

CPU: compiles against the stub boundary types, links the C ABI library and - with no GPU - both
calls return false with the library's error text (no CPU fallback).
GPU: its states / bounds / truncation equal the Python binding's on the same inputs (bit for bit:
same library, same kernels) and the oracle's within the bounds tolerance."""
import os
import subprocess

import numpy as np
import pytest

from path_optimizer_2_b200 import sharedmap, solver
from tests.test_bounds import pinch_map, straight_line

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "frontend_test")


def _build():
    solver.build_library()
    src = os.path.join(ROOT, "tests", "cpp", "frontend_test.cpp")
    deps = [src, os.path.join(ROOT, "include", "pqp_reference_path.hpp"), os.path.join(ROOT, "include", "pqp_bounds.h"),
            solver.LIB_PATH]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-o", BIN, src, solver.LIB_PATH,
                               "-Wl,-rpath," + os.path.dirname(solver.LIB_PATH)])
    return BIN


def _write(path, dm, rows, max_s):
    with open(path, "w") as f:
        f.write("%d %d %.17g\n" % (dm.rows, dm.cols, dm.res))
        np.savetxt(f, dm.dist.reshape(1, -1), fmt="%.9g")
        f.write("%d\n" % rows.shape[1])
        np.savetxt(f, rows, fmt="%.17g")
        f.write("%.17g\n" % max_s)


def _run(exe, path):
    out = subprocess.run([exe, path], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    table = np.array([[float(v) for v in ln.split()] for ln in out[2:]]) if len(out) > 2 else np.zeros((0, 13))
    return out[0], out[1], table


def test_frontend_dropin_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = _build()
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    dm = pinch_map(gap=3.4)
    lb = straight_line(dm, 50, start_y=22.0, length=40.0)
    p = str(tmp_path / "fe.txt")
    _write(p, dm, lb.spline_rows(0), 40.0)
    l0, l1, table = _run(exe, p)
    assert l0.startswith("states 0 0") and l1.startswith("bounds 0 0 0") and len(table) == 0


@pytest.mark.gpu
def test_frontend_dropin_matches_python_binding(tmp_path):
    from oracle import bounds_oracle as bo
    from path_optimizer_2_b200 import bounds
    exe = _build()
    dm = sharedmap.DistanceMap()
    ln = sharedmap.make_lines(2, 120, dmap=dm)
    pbn = bounds.PathBounds(dm.dist, dm.res)
    for b in range(2):
        rows = ln.spline_rows(b)
        max_s = float(rows[0, -1])
        p = str(tmp_path / ("fe%d.txt" % b))
        _write(p, dm, rows, max_s)
        l0, l1, table = _run(exe, p)
        os_, ox, oy, oh, ok = bo.build_states(rows, max_s)
        n = len(os_)
        assert l0.split()[:3] == ["states", "1", str(n)] and l1.split()[:4] == ["bounds", "1", str(n), "0"]
        assert np.max(np.abs(table[:, 0:5] - np.stack((os_, ox, oy, oh, ok), axis=1))) < 1e-9
        st, cv, nn, total = pbn.build_states(rows[None], [rows.shape[1]], [max_s], n)
        assert nn[0] == n and np.array_equal(table[:, 0:4], st[0].T) and np.array_equal(table[:, 4], cv[0])
        gb, gnv = pbn.compute(st, nn, rows[None], [rows.shape[1]])
        assert gnv[0] == n and np.array_equal(table[:, 5:11], gb[0].T)
        # the stored circle centres (VehicleStateBound::SingleBound::set, data_struct.hpp:82-88)
        assert np.allclose(table[:, 11], table[:, 1] + 3.9 * np.cos(table[:, 3]), atol=1e-12)
    pbn.close()
    # a blocked corridor: states are cut at the first blocked one, whose bound is kept aside
    dmp = pinch_map(gap=1.6)
    lb = straight_line(dmp, 100, start_y=22.0, length=40.0)
    p = str(tmp_path / "pinch.txt")
    _write(p, dmp, lb.spline_rows(0), 40.0)
    l0, l1, table = _run(exe, p)
    ob, onv = bo.update_bounds(dmp.dist, dmp.res, lb.spline_rows(0), *lb.states[0])
    assert l1.split()[:4] == ["bounds", "1", str(onv), "1"] and len(table) == onv
    assert np.max(np.abs(table[:, 5:11] - ob[:, :onv].T)) < 1e-9
