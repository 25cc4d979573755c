#!/bin/bash
# a variant of the WHOLE library (every translation unit rebuilt with extra flags): tools/build_variant_all.sh NAME "-D..."
set -e
cd "$(dirname "$0")/.."
out=jaxoplanet_b200/lib/variants; mkdir -p $out/$1.d
name=$1; shift
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --fmad=true $*"
for src in jxp_kernels jxp_partials jxp_transit jxp_walk jxp_state jxp_peer; do
  nvcc $F -c jaxoplanet_b200/csrc/$src.cu -o $out/$name.d/$src.o &
done
wait
nvcc -shared -o $out/$name.so $out/$name.d/*.o -gencode arch=compute_100a,code=sm_100a
rm -rf $out/$name.d
echo built $out/$name.so
