#!/bin/bash
# build_variant.sh <name> <source basename without .cu> "<extra nvcc flags>": relinks libs2c_b200 with one translation unit
# recompiled under extra flags -> build/variants/lib_<name>.so (A/B runs: S2C_B200_LIB=build/variants/lib_<name>.so)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $3 -c zk_symmetric_crypto_b200/csrc/$2.cu -o build/variants/$2_$1.o
objs=$(ls build/*.o | grep -v "/$2.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$1.so $objs build/variants/$2_$1.o -lcudart -ldl
echo built build/variants/lib_$1.so
