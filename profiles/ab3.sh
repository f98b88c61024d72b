mkdir -p gpurun_out/r2
run() { python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary "$@" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1:], round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4))" "$@"; }
echo base
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -q -x 2>&1 | tail -2
run --inflight 1; run; run --n 120; run --n 400
cp path_optimizer_2_b200/libpqp_b200_variant.so path_optimizer_2_b200/libpqp_b200.so
echo variant
run --inflight 1; run; run --n 120
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
