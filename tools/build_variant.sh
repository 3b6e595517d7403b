#!/bin/bash
# usage: tools/build_variant.sh NAME "<extra nvcc flags for klt.cu>"  ->  build/variants/libcoslam_NAME.so
# (kernel tuning experiments; select at run time with COSLAM_B200_LIB=build/variants/libcoslam_NAME.so)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
N=$1; shift
C=coslam_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC --expt-relaxed-constexpr -prec-sqrt=false -prec-div=false -Xptxas -v "$@" \
  -c $C/klt.cu -o build/variants/klt_$N.o 2> build/variants/klt_$N.log
grep -A2 "klt_front\|klt_gain_fused" build/variants/klt_$N.log | grep -E "spill|Used" || true
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libcoslam_$N.so \
  $C/common.o build/variants/klt_$N.o $C/pose.o $C/ba.o $C/posegraph.o -ldl
