for f in 0 16 32 48 64 112 8; do echo "flags=$f"; CM_DEBUG_FLAGS=$f timeout 100 python tools/quick_mlp.py 2>&1 | head -1; done
