#!/bin/bash
# Rebuild gemm.cu with each ablation knob (SA_X_*, see gemm.cu header) and time one shape:
#   tools/gemm_ablate.sh M K N [geglu] [tile]        (restores the product build at the end)
set -e
cd "$(dirname "$0")/.."
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -I include"
OBJ=stableanimator_b200/csrc/_obj
for v in "" NOEPI LDONLY NOSTORE; do
  $NV ${v:+-DSA_X_$v} -c stableanimator_b200/csrc/gemm.cu -o $OBJ/gemm.o
  nvcc -shared -o stableanimator_b200/libsa_b200.so $OBJ/*.o
  echo "== variant: ${v:-full}"
  python tools/gemm_one.py "$@"
done
$NV -c stableanimator_b200/csrc/gemm.cu -o $OBJ/gemm.o
nvcc -shared -o stableanimator_b200/libsa_b200.so $OBJ/*.o
