#!/bin/bash
# GPU call 1 of this session: full GPU test suite, MLP stage-isolation timings, source-level ncu of mm2.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/c1_pytest.txt
cat gpurun_out/c1_pytest.txt
for f in 0 1 2 4 8 16 3 9; do echo "flags=$f"; CM_DEBUG_FLAGS=$f timeout 100 python tools/quick_mlp.py 4608 12288 3840 2>&1 | head -1; done > gpurun_out/c1_mlp_flags.txt 2>&1
cat gpurun_out/c1_mlp_flags.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:mlp_kernel -s 12 -c 2 -o gpurun_out/c1_mlp -f python tools/quick_mlp.py 4608 12288 3840 > gpurun_out/c1_ncu.log 2>&1
tail -3 gpurun_out/c1_ncu.log
timeout 200 python tools/quick_attn.py 4608 784 1 24 20 2>&1 | tee gpurun_out/c1_attn.txt
timeout 200 python tools/quick_attn.py 119056 8320 1 24 3 2>&1 | tee -a gpurun_out/c1_attn.txt
ls -la gpurun_out
