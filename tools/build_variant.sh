#!/bin/bash
# development helper: build_variants/libmtt_<name>.so with only the headline density
# instantiation (MTT_K1_PROBE) and extra -D flags;   tools/build_variant.sh name -DFOO=1 ...
set -e
name=$1; shift
root=/root/repo; csrc=$root/multitypetree_b200/csrc
mkdir -p $root/build_variants
nvcc -DMTT_K1_PROBE "$@" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 \
  --expt-relaxed-constexpr -diag-suppress 177 -I $root/include -c $csrc/k1_density.cu -o /tmp/k1_$name.o -Xptxas -v 2>&1 \
  | grep -A2 "k1_density_warp" | grep -E "spill|registers" || true
objs=$(ls $csrc/_build/*.o | grep -v k1_density)
nvcc -shared -o $root/build_variants/libmtt_$name.so $objs /tmp/k1_$name.o -gencode arch=compute_100a,code=sm_100a -lcudart_static -lpthread -ldl -lrt
