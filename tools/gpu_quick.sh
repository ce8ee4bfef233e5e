#!/bin/bash
# Development loop on the GPU box (run through gpurun): GPU tests of the three kernel families + quick timings.
#   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 100 python tools/quick_mlp.py 4608 12288 3840 2>&1 | head -1
timeout 100 python tools/quick_mlp.py 16384 12288 3840 2>&1 | head -1
timeout 200 python tools/quick_attn.py 4608 784 1 24 20 2>&1 | head -2
timeout 200 python tools/quick_attn.py 16384 2944 1 24 10 2>&1 | head -1
timeout 200 python tools/quick_attn.py 119056 8320 1 24 4 2>&1 | head -2
