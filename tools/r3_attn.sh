#!/bin/bash
# same-box check of a column-sparse attention build: parity tests, then A/B of the current build against variants
mkdir -p gpurun_out
timeout -k 10 500 python -m pytest tests -m gpu -x -q -k "attn or parity or module or csp" 2>&1 | tail -3
for i in 1 2; do
  for v in "$@"; do
    CHIPMUNK_B200_LIB=chipmunk_b200/_variants/lib_$v.so timeout -k 10 200 python tools/ab_attn.py 2>&1 | tail -3 | grep -v "N= 16384"
  done
  timeout -k 10 200 python tools/ab_attn.py 2>&1 | tail -3 | grep -v "N= 16384"
done
