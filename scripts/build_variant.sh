#!/bin/bash
# usage: scripts/build_variant.sh NAME "-DFLAG=1 ..."   -> sac_rcbf_b200/variants/librcbf_NAME.so (A/B builds for scripts/gpu_ab.sh)
set -e
NAME=$1; FLAGS=$2
D=sac_rcbf_b200/variants; mkdir -p $D/obj_$NAME
for f in rcbf_kernels rcbf_safe_unicycle rcbf_safe2_unicycle rcbf_safe_cars rcbf_gp rcbf_general; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $FLAGS -c sac_rcbf_b200/csrc/$f.cu -o $D/obj_$NAME/$f.o &
done
wait
nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a $D/obj_$NAME/*.o -o $D/librcbf_$NAME.so
rm -rf $D/obj_$NAME
echo built $D/librcbf_$NAME.so
