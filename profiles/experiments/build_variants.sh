#!/bin/bash
# Alternative builds of libndsb200.so for kernel experiments: only inst_hot.cu is recompiled with the given -D flags.
#   profiles/experiments/build_variants.sh <tag> [-DFLAG ...]   ->  ndsimulator_b200/csrc/exp/libndsb200_<tag>.so
#   NDS_EXP_UNIT=inst_grid (default inst_hot) picks the translation unit that is recompiled.
# Run with NDSB200_LIB=<that file>.
set -e
cd "$(dirname "$0")/../../ndsimulator_b200/csrc"
tag=$1; shift
unit=${NDS_EXP_UNIT:-inst_hot}
mkdir -p exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
     --expt-relaxed-constexpr -Xptxas -v "$@" -c $unit.cu -o exp/${unit}_$tag.o 2> exp/${unit}_$tag.ptxas.log
objs=$(ls *.o | grep -v "^$unit.o\$")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o exp/libndsb200_$tag.so $objs exp/${unit}_$tag.o -Xcompiler -fvisibility=hidden -ldl
grep -E "registers" exp/${unit}_$tag.ptxas.log | head -3
