#!/bin/bash
# Variant library with extra -D flags applied to several translation units:
#   scripts/build_flags.sh <name> "<flags>" <tu> [<tu> ...]      e.g.  scripts/build_flags.sh nomax "-DA2_NOMAX" attn_tc2 bd_engine
# -> scripts/_bin/lib_<name>.so (select with BD_LIB_PATH).  Needs the regular build's objects in boxdreamer_b200/_build.
set -e
name=$1; flags=$2; shift 2
B=boxdreamer_b200/_build; C=boxdreamer_b200/csrc
mkdir -p scripts/_bin
BASE="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a"
objs=""
for o in $B/*.o; do
  n=$(basename $o .o); hit=0
  for t in "$@"; do [ "$t" == "$n" ] && hit=1; done
  if [ $hit == 1 ]; then nvcc $BASE $flags -c $C/$n.cu -o scripts/_bin/${name}_$n.o; objs="$objs scripts/_bin/${name}_$n.o"; else objs="$objs $o"; fi
done
nvcc -shared -o scripts/_bin/lib_${name}.so $objs -gencode arch=compute_100a,code=sm_100a -cudart static -Xcompiler -fPIC
echo scripts/_bin/lib_${name}.so
