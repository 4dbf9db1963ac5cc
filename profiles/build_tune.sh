#!/bin/bash
# build a tuning variant of libse3mpc.so: profiles/build_tune.sh <name> [nvcc -D flags...]  ->  build/tune/lib_<name>.so
set -e
name=$1; shift
mkdir -p build/tune
cd learningagileflight_se3_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC -I../../include "$@" \
    -o ../../build/tune/lib_$name.so se3mpc.cu
