#!/bin/bash
# same-box check of a dense-kernel build: correctness first, then A/B against chipmunk_b200/_variants/lib_base.so
mkdir -p gpurun_out
timeout -k 10 200 python tools/quick_dense.py > gpurun_out/r3_quick_dense.txt 2>&1; echo "quick_dense rc=$? fails=$(grep -c FAIL gpurun_out/r3_quick_dense.txt)"
timeout -k 10 400 python -m pytest tests -m gpu -x -q -k "dense or colsum or module or select or full" 2>&1 | tail -2
for i in 1 2; do
  CHIPMUNK_B200_LIB=chipmunk_b200/_variants/lib_base.so timeout -k 10 120 python tools/ab_dense16k.py 2>&1 | tail -1
  timeout -k 10 120 python tools/ab_dense16k.py 2>&1 | tail -1
done
for v in "$@"; do
  CHIPMUNK_B200_LIB=chipmunk_b200/_variants/lib_$v.so timeout -k 10 120 python tools/ab_dense16k.py 2>&1 | tail -1
done
CHIPMUNK_B200_LIB=chipmunk_b200/_variants/lib_base.so timeout -k 10 200 python tools/ab_dense_c3.py 2>&1 | tail -1
timeout -k 10 200 python tools/ab_dense_c3.py 2>&1 | tail -1
