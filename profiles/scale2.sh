#!/bin/bash
# 2-GPU sanity of both multi-GPU workloads after the last kernel changes
mkdir -p gpurun_out
for wl in cold receding; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --workload $wl > gpurun_out/scale2_$wl.json 2> gpurun_out/scale2_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale2_$wl.json").read().strip().splitlines()[-1])
    print("$wl N=2 value %.0f e2e %.0f ms/step %.3f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
except Exception as e:
    print("$wl FAILED", e); print(open("gpurun_out/scale2_$wl.err").read()[-1500:])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-200
