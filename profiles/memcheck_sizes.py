"""Small solves at every chunk size (C = 1 .. 16) of both storage policies, for a compute-sanitizer memcheck pass."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from path_optimizer_2_b200 import abi, solver, synthetic

for n in (20, 60, 120, 240, 400):
    for bits in (0, 16, 2):  # default policy, shared memory forced, FP64 instantiation
        hb = synthetic.make_batch(3, 9, n)
        sv = solver.PathQpSolver(abi.default_params(reserved=bits), n_max=n, batch_max=hb.batch)
        r = sv.solve(hb, full=True)
        r2 = sv.resolve(hb.with_linearisation(r.sol), full=True)
        print(n, bits, r.status.tolist(), r2.iters.tolist())
        sv.close()
print("done")
