"""Small driver for ncu captures: a few solves of one batch through the C ABI."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from path_optimizer_2_b200 import abi, solver, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=444)
ap.add_argument("--n", type=int, default=240)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--warm", action="store_true")
ap.add_argument("--option-bits", type=int, default=0)
a = ap.parse_args()
hb = synthetic.make_batch(3, a.batch, a.n)
sv = solver.PathQpSolver(abi.default_params(reserved=a.option_bits), n_max=a.n, batch_max=a.batch)
for _ in range(a.reps):
    res = sv.solve(hb)
    if a.warm:
        sv.resolve(hb.with_linearisation(res.sol))
print("kernel ms", sv.last_kernel_ms, "mean iters", res.iters.mean(), "info", sv.kernel_info)
