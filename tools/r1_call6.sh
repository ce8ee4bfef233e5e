#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 100 python tools/quick_mlp.py 4608 12288 3840 2>&1 | head -1
for f in 0 5; do echo "flags=$f"; CM_DEBUG_FLAGS=$f timeout 100 python tools/quick_attn.py 16384 2944 1 24 10 2>&1 | head -1; done
timeout 200 python tools/quick_attn.py 4608 784 1 24 20 2>&1 | head -1
timeout 200 python tools/quick_attn.py 119056 8320 1 24 4 2>&1 | head -1
timeout 200 python tools/quick_dense.py 2>&1 | tail -3
