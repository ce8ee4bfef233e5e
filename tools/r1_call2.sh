#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp_gpu.py tests/test_attn_gpu.py tests/test_modules_gpu.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/c2_pytest.txt
for f in 0 8 1; do echo "flags=$f"; CM_DEBUG_FLAGS=$f timeout 100 python tools/quick_mlp.py 4608 12288 3840 2>&1 | head -1; done 2>&1 | tee gpurun_out/c2_mlp.txt
timeout 200 python tools/quick_attn.py 4608 784 1 24 20 2>&1 | tee gpurun_out/c2_attn.txt
timeout 200 python tools/quick_attn.py 119056 8320 1 24 3 2>&1 | tee -a gpurun_out/c2_attn.txt
