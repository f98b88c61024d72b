#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, the reference arm, the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "reference rc=$?"
timeout 600 python bench.py --gpus 1 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python - <<PY
import json
r=json.loads(open("gpurun_out/final_reference.json").read().strip().splitlines()[-1])
d=json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print("reference: %.0f %s on %d cores (%s)" % (r["value"], r["unit"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["kind"]))
print("b200: value %.0f e2e %.0f ms/step %.3f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["traffic"])
print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
