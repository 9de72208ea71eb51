#!/bin/bash
# build a variant of libivs.so with extra nvcc flags for some sources (development aid, see tools/ab.sh):
#   [IVS_VARIANT_SRC="ivs_stream8 ivs_track"] tools/mkvariant.sh NAME [-DIVS_LAG=2 ...]
#   -> inovesa_b200/lib/variants/libivs_NAME.so
set -e
name=$1; shift
cd "$(dirname "$0")/.."
python -m inovesa_b200.build > /dev/null
mkdir -p inovesa_b200/lib/variants /tmp/ivs_variants
srcs=${IVS_VARIANT_SRC:-ivs_fused}
objs=""; skip=""
for src in $srcs; do
  obj=/tmp/ivs_variants/${src}_$name.o
  nvcc -ccbin /usr/bin/g++ -std=c++17 -O3 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a \
    -Xcompiler -fPIC,-ffp-contract=off "$@" -c inovesa_b200/csrc/$src.cu -o $obj
  objs="$objs $obj"; skip="$skip -e /$src.cu.o"
done
nvcc -ccbin /usr/bin/g++ -shared -gencode arch=compute_100a,code=sm_100a -o inovesa_b200/lib/variants/libivs_$name.so \
  $objs $(ls inovesa_b200/build/*.o | grep -v $skip) -lcudart_static -ldl -lrt -lpthread
echo inovesa_b200/lib/variants/libivs_$name.so
