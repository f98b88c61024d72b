mkdir -p gpurun_out/r2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2/scale8_numa.json 2> gpurun_out/r2/scale8_numa.err
python -c "
import json
d=json.loads(open('gpurun_out/r2/scale8_numa.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), d['e2e'].get('host_numa_node'), d['ms_per_step'], d['config']['secondary']['configs[4]']['value'])"
tail -3 gpurun_out/r2/scale8_numa.err
