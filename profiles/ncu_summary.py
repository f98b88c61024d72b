"""Summarise an .ncu-rep (raw page + source page) into a short text report."""
import csv
import subprocess
import sys

import numpy as np

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sass__inst_executed_local_loads",
        "sm__cycles_elapsed.max", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
for i, h in enumerate(hdr):
    if h in want:
        print("%-80s %-14s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
sh = srows[1]
ix = {h: i for i, h in enumerate(sh)}
stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
ex = []
for r in srows[2:]:
    if len(r) < len(sh):
        continue
    for s in stalls:
        try:
            tot[s] += int(r[ix[s]])
        except ValueError:
            pass
    try:
        ex.append(int(r[ix["Instructions Executed"]]))
    except ValueError:
        ex.append(0)
ex = np.array(ex)
tsum = max(1, sum(tot.values()))
print("SASS instructions: %d, executed at least once: %d" % (len(ex), int((ex > 0).sum())))
thr = ex.max() / 4 if len(ex) else 0
print("hot instructions (>= 1/4 of max exec count): %d covering %.1f %% of executed instructions" % (
    int((ex >= thr).sum()), 100.0 * ex[ex >= thr].sum() / max(1, ex.sum())))
print("warp stall samples: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tsum)
                                         for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]))
