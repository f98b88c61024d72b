#!/bin/bash
# Round-2 evidence in one GPU call (run from the repo root on the GPU box):
#   1. launch list of the default bench command (gpu__time_duration per launch);
#   2. one --set full capture of the ADMM kernel on the full 8192 x 240 batch, default (cold-only) handle
#      and a handle that keeps its warm state, summarised + DRAM traffic records for bench.py;
#   3. SASS evidence of the built library (tcgen05.ld/st = LDTM/STTM, cp.async.bulk = UBLKCP, mbarrier = SYNCS,
#      packed FP32 = FFMA2, no MMA).
# Outputs go to gpurun_out/r2/ (scratch); copy what is to be judged into profiles/r2/.
set -u
OUT=gpurun_out/r2
mkdir -p $OUT
TAG=${1:-default}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > $OUT/bench_under_ncu_$TAG.log 2>&1
for BITS in 128 0; do
  ncu --set full --clock-control none --import-source on -k regex:pqp_admm -s 2 -c 1 -f -o $OUT/admm_full_${TAG}_bits$BITS \
      python profiles/prof_solve.py --batch 8192 --n 240 --reps 3 --option-bits $BITS > $OUT/prof_${TAG}_bits$BITS.log 2>&1
  python profiles/ncu_summary.py $OUT/admm_full_${TAG}_bits$BITS.ncu-rep > $OUT/ncu_${TAG}_bits${BITS}_summary.txt 2>&1
done
python profiles/traffic_from_ncu.py $OUT/admm_full_${TAG}_bits128.ncu-rep 8192 240 $OUT/traffic_cold_n240_coldonly.json \
    "cold-only handle (option bit 128): inputs once + solution once; Ruiz / delta_y scratch is per resident warp and stays in L2"
python profiles/traffic_from_ncu.py $OUT/admm_full_${TAG}_bits0.ncu-rep 8192 240 $OUT/traffic_cold_n240.json \
    "handle that keeps the warm state for pqp_resolve: + 18 KB per instance of scaled x, z, y"
python profiles/ncu_lines.py $OUT/admm_full_${TAG}_bits128.ncu-rep path_optimizer_2_b200/csrc/pqp_kernel.cuh 25 > $OUT/ncu_${TAG}_functions.txt 2>&1
{
  echo "# cuobjdump -sass path_optimizer_2_b200/libpqp_b200.so | mnemonic counts"
  cuobjdump -sass path_optimizer_2_b200/libpqp_b200.so > $OUT/sass_full.txt
  for m in LDTM STTM UBLKCP SYNCS FFMA2 FMUL2 FADD2 HMMA UTCMMA QMMA IMMA DFMA FFMA; do
    printf "%-8s %s\n" $m $(grep -c "\b$m" $OUT/sass_full.txt)
  done
  echo "# first occurrences"
  grep -m3 "LDTM" $OUT/sass_full.txt; grep -m3 "STTM" $OUT/sass_full.txt; grep -m2 "UBLKCP" $OUT/sass_full.txt; grep -m2 "FFMA2" $OUT/sass_full.txt
} > $OUT/sass_excerpt_$TAG.txt
rm -f $OUT/sass_full.txt
ls -la $OUT
