#!/bin/bash
# Builds a PROFILING copy of the library with per-phase clock counters into a scratch dir and prints the breakdown.
# The production library (flemingo_b200/_lib) is not touched: the scratch copy is selected by tools/phase_clocks.py.
set -e
out=${1:-tools}
mkdir -p "$out/_phase_lib"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC \
     -DFL_PHASE_CLOCKS -I include -o "$out/_phase_lib/libflemingo_b200.so" flemingo_b200/csrc/flemingo_b200.cu flemingo_b200/csrc/builders.cu
