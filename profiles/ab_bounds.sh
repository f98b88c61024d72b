mkdir -p gpurun_out/r2
t() { python bench.py --workload sharedmap --batch 1024 --n 120 --steps 5 --warmup 3 --no-cpu --no-secondary 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']; print('$1', c['bounds_kernel']['ms_per_batch'], c['bounds_kernel']['max_abs_diff_vs_oracle'], c['plan_pipeline']['ms_per_batch'])"; }
t fp32_fast_path; t fp32_fast_path
timeout 300 python -m pytest tests/test_bounds.py tests/test_frontend_dropin.py -m gpu -q 2>&1 | tail -2
cp path_optimizer_2_b200/libpqp_b200_variant.so path_optimizer_2_b200/libpqp_b200.so
t fp32_fast_path_72regs; t fp32_fast_path_72regs
