#!/bin/bash
# A/B builds of one translation unit with extra -D flags: scripts/build_variant.sh <name> <file.cu> [-DFOO=1 ...]
# -> scripts/_bin/lib_<name>.so (select with BD_LIB_PATH).  Needs the regular build's objects in boxdreamer_b200/_build.
set -e
name=$1; src=$2; shift 2
B=boxdreamer_b200/_build; C=boxdreamer_b200/csrc
mkdir -p scripts/_bin
nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a "$@" -c $C/$src -o scripts/_bin/${name}_${src%.cu}.o
objs=""
for o in $B/*.o; do
  if [ "$(basename $o)" == "${src%.cu}.o" ]; then objs="$objs scripts/_bin/${name}_${src%.cu}.o"; else objs="$objs $o"; fi
done
nvcc -shared -o scripts/_bin/lib_${name}.so $objs -gencode arch=compute_100a,code=sm_100a -cudart static -Xcompiler -fPIC
echo scripts/_bin/lib_${name}.so
