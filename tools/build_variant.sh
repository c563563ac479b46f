#!/usr/bin/env bash
# Build a tuning variant of libnm_b200.so: tools/build_variant.sh <name> <extra nvcc flags for nm_pyramid.cu...>
# -> build/variants/libnm_b200_<name>.so (same objects as the product, nm_pyramid.cu recompiled with the flags).
set -e
name=$1; shift
mkdir -p build/variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude"
$NV "$@" -c niftymatch_b200/csrc/nm_pyramid.cu -o build/variants/nm_pyramid_$name.o
objs=$(ls build/obj/*.o | grep -v nm_pyramid.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libnm_b200_$name.so build/variants/nm_pyramid_$name.o $objs
rm -f build/variants/nm_pyramid_$name.o
ls -la build/variants/libnm_b200_$name.so
