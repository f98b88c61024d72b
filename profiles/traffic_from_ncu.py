"""Write profiles/r2/traffic_*.json from an .ncu-rep (--set full capture of the ADMM kernel) so that
bench.py's roofline.traffic always comes from a capture of the SAME round's kernel.
usage: traffic_from_ncu.py <report.ncu-rep> <batch> <n> <out.json> [note]"""
import csv
import json
import subprocess
import sys

rep, batch, n, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
note = sys.argv[5] if len(sys.argv) > 5 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def get(name):
    v = float(vals[col[name]].replace(",", ""))
    u = units[col[name]].lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    return v * scale


rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
rec = {"kernel": vals[col["Kernel Name"]] if "Kernel Name" in col else "pqp_admm_kernel_tmem", "batch": batch, "n": n,
       "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
       "dram_bytes_per_instance": (rd + wr) / batch, "algorithmic_bytes_per_instance": 104 * n + 56,
       "gpu_time_ms": get("gpu__time_duration.sum") / 1e6 if units[col["gpu__time_duration.sum"]].lower() in ("ns", "nsecond") else get("gpu__time_duration.sum"),
       "warp_instructions": get("smsp__inst_executed.sum") if "smsp__inst_executed.sum" in col else None,
       "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active") if "smsp__issue_active.avg.pct_of_peak_sustained_active" in col else None,
       "registers_per_thread": get("launch__registers_per_thread") if "launch__registers_per_thread" in col else None,
       "note": note, "source": "ncu --set full --clock-control none (profiles/capture_r2.sh), report " + rep.split("/")[-1]}
with open(out, "w") as f:
    json.dump(rec, f, indent=1)
print(json.dumps(rec))
