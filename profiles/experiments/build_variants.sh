#!/bin/bash
# Alternative builds of libndsb200.so for kernel experiments: only inst_hot.cu is recompiled with the given -D flags.
#   profiles/experiments/build_variants.sh <tag> [-DFLAG ...]   ->  ndsimulator_b200/csrc/exp/libndsb200_<tag>.so
# Run with NDSB200_LIB=<that file>.
set -e
cd "$(dirname "$0")/../../ndsimulator_b200/csrc"
tag=$1; shift
mkdir -p exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
     --expt-relaxed-constexpr -Xptxas -v "$@" -c inst_hot.cu -o exp/inst_hot_$tag.o 2> exp/inst_hot_$tag.ptxas.log
objs=$(ls *.o | grep -v '^inst_hot.o$')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o exp/libndsb200_$tag.so $objs exp/inst_hot_$tag.o -Xcompiler -fvisibility=hidden -ldl
grep -A1 "Li5ELi1ELi10ELi10ELi1ELi9E" exp/inst_hot_$tag.ptxas.log | grep -E "registers|spill" | head -4
