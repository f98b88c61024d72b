#!/bin/bash
# one --set full capture of the ADMM kernel (full 8192 x 240 batch, cold-only handle) + summaries; usage: capture_one.sh TAG
set -u
OUT=gpurun_out/r2
mkdir -p $OUT
TAG=${1:-one}
ncu --set full --clock-control none --import-source on -k regex:pqp_admm -s 2 -c 1 -f -o $OUT/admm_full_${TAG} \
    python profiles/prof_solve.py --batch 8192 --n 240 --reps 3 --option-bits 128 > $OUT/prof_${TAG}.log 2>&1
python profiles/ncu_summary.py $OUT/admm_full_${TAG}.ncu-rep > $OUT/ncu_${TAG}_summary.txt 2>&1
python profiles/ncu_lines.py $OUT/admm_full_${TAG}.ncu-rep path_optimizer_2_b200/csrc/pqp_kernel.cuh 30 > $OUT/ncu_${TAG}_functions.txt 2>&1
grep -E "time_duration|inst_executed.sum|issue_active|warp stall|SASS" $OUT/ncu_${TAG}_summary.txt
