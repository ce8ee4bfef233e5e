for f in 0 5 13 7 15 2; do echo "flags=$f"; CM_DEBUG_FLAGS=$f timeout 100 python tools/quick_mlp.py 2>&1 | head -1; done
