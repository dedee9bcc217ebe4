#!/bin/bash
# Builds an alternative libwrfgpu for tuning experiments (selected at run time with WRFGPU_LIB):
#   [FMAD=true] bash variant.sh <name> [-DMACRO=VALUE ...]   ->  ../variants/libwrfgpu_<name>.so
# FMAD=true lets nvcc contract a*b+c (the bit-exact build forbids it; profiles/README.md has the A/B)
set -e
cd "$(dirname "$0")"
name=$1; shift
out=../variants; mkdir -p $out /tmp/wgv_$name
for f in solver bc prep tend tend_tile advect_fast mom_fast pd_fast diffusion deform diffusion2 acoustic ieva scalar moist bdy halo history decomp; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=${FMAD:-false} -Xcompiler -fPIC "$@" -c $f.cu -o /tmp/wgv_$name/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libwrfgpu_$name.so /tmp/wgv_$name/*.o -lcudart
echo built $out/libwrfgpu_$name.so
