#!/bin/bash
# build and run a tests/probes/*.cu hardware probe on the GPU box:  tools/r3_probe.sh probe_mma_power [more ...]
mkdir -p gpurun_out
for p in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/$p tests/probes/$p.cu 2>/dev/null
  (nvidia-smi --query-gpu=power.draw,clocks.sm,clocks_throttle_reasons.active -lms 200 --format=csv,noheader > gpurun_out/r3_${p}_smi.txt &) 
  timeout -k 5 120 /tmp/$p | tee gpurun_out/r3_$p.txt
done
sort gpurun_out/r3_${p}_smi.txt | uniq -c | sort -rn | head -12
