#!/bin/bash
# e2e throughput of the streamed host call vs. its minimum chunk size
for mc in 256 512 1024; do
  echo "PQP_STREAM_MIN_CHUNK=$mc"
  PQP_STREAM_MIN_CHUNK=$mc python profiles/sweep.py "--n 120 --batch 1024 --steps 10" "--n 120 --batch 2048 --steps 10" "--n 120 --batch 4096 --steps 10" "--n 240 --batch 2048 --steps 10" "--n 240 --batch 8192"
done
