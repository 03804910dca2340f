#!/bin/bash
# Profiling builds that drop one phase of place_fast_kernel (results are wrong; timing only).
set -e
for v in SKIP_REFINE SKIP_DP; do
  mkdir -p tools/_phase_lib/$v
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC \
       -DFL_$v -I include -o tools/_phase_lib/$v/libflemingo_b200.so flemingo_b200/csrc/flemingo_b200.cu &
done
wait
