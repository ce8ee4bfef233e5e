#!/bin/bash
# Build a variant of the library for same-box A/B timing: one source recompiled with extra flags, linked with the
# objects of the regular build.   tools/build_variant.sh <tag> <source.cu> <nvcc flags...>   -> chipmunk_b200/_variants/lib_<tag>.so
set -e
tag=$1; src=$2; shift 2
cd "$(dirname "$0")/../chipmunk_b200"
mkdir -p _variants
base=$(basename "$src" .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c csrc/$base.cu -o _build/${base}_$tag.o
objs=$(ls _build/*.o | grep -v "_build/${base}\(_[A-Za-z0-9]*\)\?\.o" | grep -v "_[a-z0-9]*\.o$" || true)
objs=$(for s in csrc/*.cu; do b=$(basename $s .cu); [ "$b" != "$base" ] && echo _build/$b.o; done)
nvcc -shared -o _variants/lib_$tag.so _build/${base}_$tag.o $objs -lcuda
echo chipmunk_b200/_variants/lib_$tag.so
