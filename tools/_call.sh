for L in P0 P1 P2 P0 P1 P2; do echo lib$L
CHIPMUNK_B200_LIB=$PWD/chipmunk_b200/lib$L.so timeout 200 python tools/quick_dense.py 2>&1 | tail -1
done
CHIPMUNK_B200_LIB=$PWD/chipmunk_b200/libP1.so timeout 300 python -m pytest tests/test_attn_gpu.py -m gpu -q -k colsum 2>&1 | tail -2
