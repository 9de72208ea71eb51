#!/bin/bash
# build a variant of libivs.so with extra nvcc flags for ivs_fused.cu (development aid, see tools/ab.sh):
#   tools/mkvariant.sh NAME [-DIVS_LAG=2 ...]   -> inovesa_b200/lib/variants/libivs_NAME.so
set -e
name=$1; shift
cd "$(dirname "$0")/.."
python -m inovesa_b200.build > /dev/null
mkdir -p inovesa_b200/lib/variants /tmp/ivs_variants
src=${IVS_VARIANT_SRC:-ivs_fused}   # IVS_VARIANT_SRC=ivs_stream8 for the column-block kernel
obj=/tmp/ivs_variants/${src}_$name.o
nvcc -ccbin /usr/bin/g++ -std=c++17 -O3 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-ffp-contract=off "$@" -c inovesa_b200/csrc/$src.cu -o $obj
nvcc -ccbin /usr/bin/g++ -shared -gencode arch=compute_100a,code=sm_100a -o inovesa_b200/lib/variants/libivs_$name.so \
  $obj $(ls inovesa_b200/build/*.o | grep -v "/$src.cu.o") -lcudart_static -ldl -lrt -lpthread
echo inovesa_b200/lib/variants/libivs_$name.so
