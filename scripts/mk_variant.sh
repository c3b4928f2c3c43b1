#!/bin/bash
# A/B build: scripts/mk_variant.sh NAME TU.cu [-DFLAG=..]...  ->  ab/lib_NAME.so (the default objects + TU recompiled with
# the extra flags).  Select it on the GPU box with RCBF_LIB_PATH=ab/lib_NAME.so (sac_rcbf_b200/_lib.py).
set -e
cd "$(dirname "$0")/.."
name=$1; tu=$2; shift 2
mkdir -p ab sac_rcbf_b200/build/ab_$name
obj=sac_rcbf_b200/build/ab_$name/${tu%.cu}.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c sac_rcbf_b200/csrc/$tu -o $obj
others=$(ls sac_rcbf_b200/build/*.o | grep -v "/${tu%.cu}.o")
nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a $others $obj -o ab/lib_$name.so
echo ab/lib_$name.so
