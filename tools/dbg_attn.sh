for f in 0 1 2 4 3 5 6 7; do echo "flags=$f"; CM_DEBUG_FLAGS=$f timeout 100 python tools/quick_attn.py 16384 2944 1 24 5 2>&1 | head -1; done
