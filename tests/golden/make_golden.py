"""Regenerates tests/golden/oracle_golden.json from the oracle (run from the repo root).
The reference ships no golden vectors for this path and its solver cannot
be built or imported here, so these vectors freeze the ORACLE's behaviour only."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle  # noqa: E402
from path_optimizer_2_b200 import abi, synthetic  # noqa: E402

cases = []
prm = abi.default_params()
for cfg, index, n in [(3, 0, 120), (3, 1, 120), (3, 5, 240), (3, 9, 240), (7, 2, 33), (103, 4, 3)]:
    k, inst, ne, _ = synthetic.make_instance(cfg, index, n)
    s = oracle.OracleSolver(prm, k, inst, ne)
    st = s.solve()
    stride = max(1, n // 12)
    cases.append(dict(cfg=cfg, index=index, n=n, status=st, iters=s.iters, cost=s.cost, stride=stride,
                      sol_sampled=s.sol()[:, ::stride].ravel().tolist()))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w") as f:
    json.dump(dict(note="oracle-produced (parity unpinned: no reference-produced vectors exist)", cases=cases), f, indent=1)
print("wrote", len(cases), "cases")
