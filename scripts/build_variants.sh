#!/bin/bash
# Build hot-loop variants of the library into circuitree_b200/variants/ (timed side by side by
# scripts/variants.sh on the GPU box).  usage: build_variants.sh name:"-DFLAG=1 ..." ...
set -e
cd "$(dirname "$0")/.."
mkdir -p circuitree_b200/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  echo "== $name: $flags"
  (cd circuitree_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
     $flags -o ../variants/lib_$name.so engine.cu native_tree.cpp) &
done
wait
ls -la circuitree_b200/variants
