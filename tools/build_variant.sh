#!/bin/bash
# Experiment helper: rebuild ONE translation unit with extra -D flags and link it with the production objects into
# speech_separation_b200/variants/libvatss_<name>.so (git-ignored; select it with VATSS_LIB_OVERRIDE=<path>).
#   tools/build_variant.sh poly2 tc_attn3 -DA3_POLY=2
set -e
name=$1; unit=$2; shift 2
cd "$(dirname "$0")/../speech_separation_b200/csrc"
mkdir -p build/var ../variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-Wall,-Wno-unused-function "$@" -c $unit.cu -o build/var/${unit}_$name.o
objs=""
for o in api frontend generic_block tail sisnr tensor_engine tc_gemm tc_lstm tc_attention tc_attn3; do
  if [ $o == $unit ]; then objs="$objs build/var/${unit}_$name.o"; else objs="$objs build/$o.o"; fi
done
/usr/local/cuda/bin/nvcc $ARCH -shared -o ../variants/libvatss_$name.so $objs -lcudart_static -lpthread -ldl -lrt
echo built ../variants/libvatss_$name.so
