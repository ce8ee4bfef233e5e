#!/bin/bash
timeout 100 python tools/quick_mlp.py 4608 12288 3840 2>&1 | head -1
timeout 100 python tools/quick_mlp.py 16384 12288 2304 2>&1 | head -1
timeout 100 python tools/quick_mlp.py 16384 12288 3840 2>&1 | head -1
