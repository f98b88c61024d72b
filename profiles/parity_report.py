#!/usr/bin/env python
"""Parity statistics of the CUDA path against the CPU oracle over whole batches (runs on the GPU box).
For every instance: status agreement, iteration-count agreement, |x_gpu - x_oracle|_inf, and - for
solved instances - OSQP's termination test evaluated in FP64 on the GPU's returned (x, y, z) against the
oracle-assembled problem. Also the FP64 instantiation (reserved bit 2), which must reproduce the oracle.
usage: python profiles/parity_report.py > profiles/r1/parity_report.txt"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from path_optimizer_2_b200 import abi, solver, synthetic  # noqa: E402


def report(label, hb, bits=0, term_sample=256):
    params = abi.default_params(reserved=bits)
    sv = solver.PathQpSolver(params, n_max=hb.n_max, batch_max=hb.batch)
    g = sv.solve(hb, full=True)
    sv.close()
    o, _ = oracle.solve_batch(abi.default_params(), hb, nthreads=oracle.max_threads(), full=True)
    B = hb.batch
    same_status = g.status == o.status
    both = (g.status == abi.PQP_SOLVED) & (o.status == abi.PQP_SOLVED)
    d_it = g.iters[both].astype(np.int64) - o.iters[both]
    nv = 6 * hb.n - 1
    dx = np.array([np.max(np.abs(g.x_full[b, :nv[b]] - o.x_full[b, :nv[b]])) for b in np.nonzero(both)[0]])
    dcost = np.abs(g.cost[both] - o.cost[both]) / np.maximum(1.0, np.abs(o.cost[both]))
    print("== %s: %d instances, n_max %d, option bits %d" % (label, B, hb.n_max, bits))
    print("   status: oracle %s | gpu %s | agree %.4f" % (np.bincount(o.status, minlength=6).tolist(),
                                                          np.bincount(g.status, minlength=6).tolist(), same_status.mean()))
    print("   iterations (both solved, %d): identical %.4f, within one check interval %.4f, mean gpu %.1f oracle %.1f, "
          "max |diff| %d" % (both.sum(), (d_it == 0).mean(), (np.abs(d_it) <= 25).mean(), g.iters[both].mean(),
                             o.iters[both].mean(), np.abs(d_it).max() if len(d_it) else 0))
    print("   |x_gpu - x_oracle|_inf: median %.3g, 99 %% %.3g, max %.3g;  relative cost difference: median %.3g, max %.3g"
          % (np.median(dx), np.quantile(dx, 0.99), dx.max(), np.median(dcost), dcost.max()))
    idx = np.nonzero(g.status == abi.PQP_SOLVED)[0]
    idx = idx[:: max(1, len(idx) // term_sample)]
    worst_p = worst_d = 0.0
    for b in idx:
        s = oracle.OracleSolver(abi.default_params(), hb.knots[b], hb.inst[b], int(hb.n[b]))
        Pd, A, l, u = s.problem()
        rep = oracle.osqp_termination_report(Pd, A, l, u, g.x_full[b, :s.nv], g.y_full[b, :s.m], g.z_full[b, :s.m])
        worst_p = max(worst_p, rep["pri_res"] / rep["eps_pri"])
        worst_d = max(worst_d, rep["dua_res"] / rep["eps_dua"])
    print("   OSQP termination test in FP64 on the returned iterates (%d sampled): worst residual / tolerance primal %.3f, "
          "dual %.3f (must be < 1 up to the kernel's FP32 norms)" % (len(idx), worst_p, worst_d))


if __name__ == "__main__":
    report("BASELINE configs[2] slice (per-instance bounds)", synthetic.make_batch(3, 2048, 240))
    report("ragged n", synthetic.make_batch(7, 1024, 240, ragged=True))
    report("n = 120", synthetic.make_batch(3, 2048, 120))
    report("n = 240, textbook form of the ADMM step (option bit 64; the default is the increment form)", synthetic.make_batch(3, 2048, 240), bits=64)
    report("n = 120, textbook form (option bit 64)", synthetic.make_batch(3, 2048, 120), bits=64)
    report("n = 60, increment form (option bit 32)", synthetic.make_batch(3, 2048, 60), bits=32)
    report("FP64 instantiation (must reproduce the oracle)", synthetic.make_batch(3, 512, 240), bits=2)
    try:
        from oracle import bounds_oracle
        from path_optimizer_2_b200 import bounds, sharedmap
        dm = sharedmap.DistanceMap()
        ln = sharedmap.make_lines(512, 120, dmap=dm)
        pbn = bounds.PathBounds(dm.dist, dm.res)
        bnd, nvv = pbn.compute(ln.states, ln.n, ln.spline, ln.k)
        pbn.close()
        worst, flips, cut = 0.0, 0, 0
        for b in range(ln.batch):
            ob, onv = bounds_oracle.update_bounds(dm.dist, dm.res, ln.spline_rows(b), *ln.states[b])
            d = np.abs(ob - bnd[b])
            flips += int((d > 1e-9).sum())
            worst = max(worst, float(d[d <= 1e-9].max()))
            cut += int(onv != nvv[b])
        print("== clearance bounds kernel vs oracle: %d paths x 120 states x 6 bounds: max |diff| %.3g, march-decision flips %d, "
              "truncation mismatches %d" % (ln.batch, worst, flips, cut))
        report("BASELINE configs[1] slice (shared map, bounds from the bounds kernel)", ln.to_host_batch(bnd, nvv))
    except Exception as e:  # noqa: BLE001
        print("shared-map part failed:", e)
