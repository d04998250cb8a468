#!/bin/bash
# Build profiling variants of the library (NAF_TC_EXP bit mask: 1 = no output stores, 2 = no q
# loads, 4 = no MMAs) into scripts/exp/ and time the C2 attention launch with each (GPU box).
set -e
cd "$(dirname "$0")/.."
for e in "$@"; do
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3,-fvisibility=hidden \
    -shared -I include -I naf_b200/csrc -DNAF_BUILDING_LIB -DNAF_WITH_TC -DNAF_TC_EXP=$e \
    -o scripts/exp/libnaf_exp$e.so naf_b200/csrc/*.cu
done
