timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_modules_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 200 python tools/quick_dense.py 2>&1 | tail -3
