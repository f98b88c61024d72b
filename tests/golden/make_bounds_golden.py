"""Regenerates tests/golden/bounds_golden.json from oracle/bounds_oracle.py (run from the repo root).
The reference ships no golden vectors for its bounds front end and cannot be built here, so these
vectors freeze the ORACLE's behaviour only (parity unpinned)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import bounds_oracle as bo  # noqa: E402
from path_optimizer_2_b200 import sharedmap  # noqa: E402

dm = sharedmap.DistanceMap()
ln = sharedmap.make_lines(3, 120, dmap=dm)
cases = []
for b in range(3):
    rows = ln.spline_rows(b)
    bounds, n_valid = bo.update_bounds(dm.dist, dm.res, rows, *ln.states[b])
    s, x, y, h, k = bo.build_states(rows, float(rows[0, -1]))
    cases.append(dict(line=b, spline=rows.tolist(), states=ln.states[b][:, ::8].tolist(), n_valid=n_valid,
                      bounds_sampled=bounds[:, ::8].tolist(), total_states=len(s),
                      states_sampled=np.stack((s, x, y, h, k))[:, ::16].tolist()))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bounds_golden.json"), "w") as f:
    json.dump(dict(note="oracle-produced over tests/golden/gridmap.png (parity unpinned: no reference-produced vectors exist)",
                   cases=cases), f)
print("wrote", len(cases), "cases")
