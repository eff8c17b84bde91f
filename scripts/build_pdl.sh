#!/bin/bash
# Variant library with programmatic dependent launch (-DBD_PDL) in the GEMM / attention / LayerNorm chain:
# scripts/_bin/lib_pdl.so, select with BD_LIB_PATH.  Not measured yet (round 2): A/B with
#   bash scripts/gpu_ab_bench.sh BD_LIB_PATH= BD_LIB_PATH=scripts/_bin/lib_pdl.so
# and run the -m gpu tests against it first (BD_LIB_PATH=... python -m pytest tests -m gpu -q).
set -e
B=boxdreamer_b200/_build; C=boxdreamer_b200/csrc
mkdir -p scripts/_bin
FLAGS="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -DBD_PDL"
objs=""
for o in $B/*.o; do
  n=$(basename $o .o)
  case $n in
    attn_tc2|gemm_tc2|kernels_simt) nvcc $FLAGS -c $C/$n.cu -o scripts/_bin/pdl_$n.o; objs="$objs scripts/_bin/pdl_$n.o";;
    *) objs="$objs $o";;
  esac
done
nvcc -shared -o scripts/_bin/lib_pdl.so $objs -gencode arch=compute_100a,code=sm_100a -cudart static -Xcompiler -fPIC
echo scripts/_bin/lib_pdl.so
