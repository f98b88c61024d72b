mkdir -p gpurun_out/r2
nvidia-smi topo -m > gpurun_out/r2/scale4_topo.txt 2>&1
for i in 0 1 2 3; do cat /sys/bus/pci/devices/$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $i | tr 'A-Z' 'a-z' | sed 's/^0000//')/numa_node; done >> gpurun_out/r2/scale4_topo.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 --no-secondary > gpurun_out/r2/scale4_numa.json 2> gpurun_out/r2/scale4_numa.err
python -c "
import json
d=json.loads(open('gpurun_out/r2/scale4_numa.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), d['e2e'].get('host_numa_node'), d['ms_per_step'])"
