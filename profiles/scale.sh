#!/bin/bash
# bench.py at N = 1, 2, 4, 8 on one box (BASELINE configs[2]/[3]) and the receding-horizon config[4] at N = 8
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 400 python bench.py --gpus 1 --no-cpu > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print("N=$n value %.0f e2e %.0f ms/step %.3f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
except Exception as e:
    print("N=$n FAILED", e); print(open("gpurun_out/scale_n$n.err").read()[-1500:])
PY
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --workload receding --steps 20 > gpurun_out/scale_receding_n8.json 2> gpurun_out/scale_receding_n8.err
tail -c 1500 gpurun_out/scale_receding_n8.json
