#!/bin/bash
# A/B of two builds of the library (PQP_LIB_PATH): usage profiles/ab_lib.sh <variant.so> "<bench args>" ...
variant=$1; shift
for spec in "$@"; do
  echo "base   : $(python profiles/sweep.py "$spec")"
  echo "variant: $(PQP_LIB_PATH=$variant python profiles/sweep.py "$spec")"
done
