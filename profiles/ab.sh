mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -q -x 2>&1 | tail -3
for K in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary --inflight $K 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($K, round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['mean_admm_iters'])"
done
for N in 120 60 400; do
python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary --n $N 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n',$N, round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['roofline']['kernel_ms'])"
done
python bench.py --workload receding --batch 512 --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('receding', round(d['value']), round(d['e2e']['value']), d['ms_per_step'])"
