#!/bin/sh
# Builds the yardstick programs (CUB sort, D2D copy) into profiles/yardsticks/build/: measurements beside the product, never
# part of it.  Run from the repository root after `python -m pbf_b200.build`.
set -e
mkdir -p profiles/yardsticks/build
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iinclude -o profiles/yardsticks/build/sort_yardstick \
    profiles/yardsticks/sort_yardstick.cu pbf_b200/libpbf_b200.so -Xlinker -rpath -Xlinker '$ORIGIN/../../../pbf_b200'
