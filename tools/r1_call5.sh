#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_modules_gpu.py -m gpu -q -x 2>&1 | tail -3
for f in 0 5; do echo "flags=$f"; CM_DEBUG_FLAGS=$f timeout 100 python tools/quick_attn.py 16384 2944 1 24 10 2>&1 | head -1; done 2>&1 | tee gpurun_out/c5_attn_flags.txt
timeout 200 python tools/quick_attn.py 4608 784 1 24 20 2>&1 | head -1
timeout 200 python tools/quick_attn.py 119056 8320 1 24 4 2>&1 | head -1
timeout 900 python tools/sweep.py --out gpurun_out/sweep.jsonl 2>&1 | tail -40
