mkdir -p gpurun_out/r2
run() { python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary "$@" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1:], round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['config'].get('warps_per_sm'))" "$@"; }
run --n 120; run --n 120 --option-bits 136; run --n 60; run --n 60 --option-bits 136; run --n 30; run --n 100 ; run --n 100 --option-bits 136; run --n 64; run --n 64 --option-bits 136
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -q -x 2>&1 | tail -2
