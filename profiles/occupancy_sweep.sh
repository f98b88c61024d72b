#!/bin/bash
# shared-memory kernel at n = 60 (18.4 KB per QP): throughput vs resident warps per SM, forced by padding
# the dynamic shared memory per CTA (PQP_SMEM_PAD). 8 = register limit (254 regs/thread).
for pad in 0 14000 20000 38000 95000; do
  echo "PQP_SMEM_PAD=$pad"
  PQP_SMEM_PAD=$pad python profiles/sweep.py "--n 60 --batch 16384 --option-bits 16"
done
for pad in 0 20000; do
  echo "n=120 PQP_SMEM_PAD=$pad"
  PQP_SMEM_PAD=$pad python profiles/sweep.py "--n 120 --batch 8192 --option-bits 16"
done
