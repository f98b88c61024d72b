#!/usr/bin/env python
"""profiles/sweep.py — run bench.py over several argument sets and print one compact line each.
usage: python profiles/sweep.py "--workload sharedmap --option-bits 0" "--n 120 --batch 1024" ..."""
import json
import os
import shlex
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for spec in sys.argv[1:]:
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu"] + shlex.split(spec)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print("%-60s value %10.0f  e2e %10.0f  kernel_ms %8.3f  iters %.1f  %s x%d" % (
            spec, d["value"], d["e2e"]["value"], d["roofline"].get("kernel_ms", 0.0),
            d["config"].get("mean_admm_iters", 0.0), d["config"].get("state_storage", "?")[:13],
            d["config"].get("warps_per_sm", 0)), flush=True)
    except Exception as e:  # noqa: BLE001
        print(spec, "FAILED", e, p.stderr[-400:], flush=True)
