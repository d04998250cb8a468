#!/bin/bash
# Build profiling variants of the library (NAF_CONV_EXP bit mask, see naf_conv_tc.cu) into
# scripts/exp/ HERE (nvcc cross-compiles), then on the GPU box:  for e in ...; do
#   NAF_B200_LIB=scripts/exp/libnaf_conv$e.so python scripts/enc_bench.py; done
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/exp
for e in "$@"; do
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3,-fvisibility=hidden \
    -shared -I include -I naf_b200/csrc -DNAF_BUILDING_LIB -DNAF_WITH_TC -DNAF_CONV_EXP=$e \
    -o scripts/exp/libnaf_conv$e.so naf_b200/csrc/*.cu &
done
wait
