#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/c3_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -2 | tee gpurun_out/c3_bench.json
