"""Attribute executed instructions and stall samples of an .ncu-rep to CUDA source lines / functions.
usage: python profiles/ncu_lines.py report.ncu-rep [kernel_header.cuh]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
files = {}
cur_file = None
per_line = defaultdict(lambda: [0, 0])  # (file, line) -> [instructions, samples]
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < 10 or not r[0]:
        continue
    try:
        ln = int(r[0])
        per_line[(cur_file, ln)][0] += int(r[hdr["Instructions Executed"]])
        per_line[(cur_file, ln)][1] += int(r[hdr["# Samples"]])
    except ValueError:
        pass
tot_i = sum(v[0] for v in per_line.values())
tot_s = sum(v[1] for v in per_line.values())
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
# map lines of the kernel header to enclosing PQP_DEV functions
if len(sys.argv) > 2:
    path = sys.argv[2]
    text = open(path).read().splitlines()
    func_at = []
    cur = "(file scope)"
    for i, l in enumerate(text, 1):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?PQP_DEV\s+[\w:<>&\s\*]+?\s+(\w+)\s*\(", l)
        if m:
            cur = m.group(1)
        func_at.append(cur)
    agg = defaultdict(lambda: [0, 0])
    for (f, ln), v in per_line.items():
        if f and f.endswith(path.split("/")[-1]) and 1 <= ln <= len(func_at):
            a = agg[func_at[ln - 1]]
        else:
            a = agg["[%s]" % (f.split("/")[-1] if f else "?")]
        a[0] += v[0]
        a[1] += v[1]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%-36s instr %5.1f %%   samples %5.1f %%" % (k, 100.0 * v[0] / tot_i, 100.0 * v[1] / max(1, tot_s)))
if len(sys.argv) > 3:
    top = sorted(per_line.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3])]
    for (f, ln), v in top:
        print("%-22s:%5d  instr %5.2f %%  samples %5.2f %%" % (f.split("/")[-1] if f else "?", ln, 100.0 * v[0] / tot_i, 100.0 * v[1] / max(1, tot_s)))
