#!/bin/bash
mkdir -p gpurun_out
echo "== variant A (CS first)"; timeout 100 python tools/quick_dense.py --time 2>&1 | tail -2 | cut -c1-200
echo "== variant B (CS late)"; CHIPMUNK_B200_LIB=chipmunk_b200/_variants/libB.so timeout 100 python tools/quick_dense.py --time 2>&1 | tail -2 | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -c 1 -f -o gpurun_out/r2_dense_cs_v2 python tools/prof_targets.py dense_cs > gpurun_out/r2_ncu_dense_cs_v2.log 2>&1
