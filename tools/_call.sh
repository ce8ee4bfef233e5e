for rep in 1 2; do for L in A B; do echo lib$L
CHIPMUNK_B200_LIB=$PWD/chipmunk_b200/lib$L.so timeout 200 python tools/quick_attn.py 16384 2944 1 24 10 2>&1 | head -1
CHIPMUNK_B200_LIB=$PWD/chipmunk_b200/lib$L.so timeout 200 python tools/quick_mlp.py 4608 12288 3840 2>&1 | head -1
done; done
