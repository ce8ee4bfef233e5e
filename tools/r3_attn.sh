#!/bin/bash
# same-box check of a column-sparse attention build: a guarded first launch, parity tests, then A/B of the current build
# against variants (every step under its own short timeout: a hung kernel must not eat the GPU budget)
mkdir -p gpurun_out
timeout -k 5 60 python tools/quick_attn.py 4608 784 1 24 5 2>&1 | head -1 || { echo "first launch hung or failed"; exit 1; }
timeout -k 5 150 python -m pytest tests -m gpu -x -q -k "attn or parity or module or csp" 2>&1 | tail -3
for i in 1 2; do
  for v in "$@"; do
    CHIPMUNK_B200_LIB=chipmunk_b200/_variants/lib_$v.so timeout -k 5 60 python tools/ab_attn.py 2>&1 | tail -3 | grep -v "N= 16384"
  done
  timeout -k 5 60 python tools/ab_attn.py 2>&1 | tail -3 | grep -v "N= 16384"
done
