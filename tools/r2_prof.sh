#!/bin/bash
# ncu captures of the round-2 kernels (run under gpurun); summaries are made here by tools/ncu_summary.py
set -x
mkdir -p gpurun_out
for t in "$@"; do
  case $t in
    select)   pat="select_columns_kernel";;
    m2i)      pat="mask_to_indices_kernel";;
    dense)    pat="dense_kernel";;
    dense_cs) pat="dense_kernel";;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -c 1 -f -o gpurun_out/r2_$t python tools/prof_targets.py $t > gpurun_out/r2_ncu_$t.log 2>&1
done
