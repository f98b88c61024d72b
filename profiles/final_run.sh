#!/bin/bash
# Round-end evidence in one GPU call: GPU test suite, the driver's bench commands (own arm + reference arm),
# ncu captures (launch list, --set full on both handle kinds, traffic, SASS excerpt), other sizes, smoke.
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/final_bench.json 2> gpurun_out/r2/final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/final_ref.json 2> gpurun_out/r2/final_ref.err
bash profiles/capture_r2.sh final > gpurun_out/r2/capture_final.log 2>&1
for N in 120 60 400; do
  python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary --n $N > gpurun_out/r2/final_n$N.json 2>/dev/null
done
for K in 1 2; do
  python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary --inflight $K 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($K, round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['roofline']['kernel_ms'])" >> gpurun_out/r2/final_inflight.log
done
python __graft_entry__.py smoke 2>&1 | tail -3
tail -c 400 gpurun_out/r2/final_bench.json
