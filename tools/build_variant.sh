#!/bin/bash
# Kernel experiments: builds variants/libsdtree_<name>.so from the same sources with extra -D flags, for
#   python bench.py --lib variants/libsdtree_<name>.so
# (the shipped library is practical_path_guiding_lab_b200/libsdtree.so, built by practical_path_guiding_lab_b200/build.py)
#   tools/build_variant.sh <name> [-DSDT_...=.. ...]
set -eu
NAME=${1:?name}; shift
cd "$(dirname "$0")/.."
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared \
  "$@" -o variants/libsdtree_${NAME}.so practical_path_guiding_lab_b200/csrc/sdtree.cu -lcudart -ldl
echo variants/libsdtree_${NAME}.so
