#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/probe_smem_bw tests/probes/probe_smem_bw.cu 2>/dev/null
timeout -k 5 60 /tmp/probe_smem_bw | tee gpurun_out/r3_probe_smem_bw.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -c 1 -f -o gpurun_out/r3_dense python tools/prof_targets.py dense > gpurun_out/r3_ncu_dense.log 2>&1
tail -2 gpurun_out/r3_ncu_dense.log
