#!/bin/bash
# a variant of the library that differs in jxp_walk.cu only: tools/build_walk_variant.sh NAME "-DJXP_WALK_N=3 ..."
set -e
cd "$(dirname "$0")/.."
out=jaxoplanet_b200/lib/variants; mkdir -p $out
name=$1; shift
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --fmad=true -Xptxas -v $*"
nvcc $F -c jaxoplanet_b200/csrc/jxp_walk.cu -o $out/$name.w.o > $out/$name.log 2>&1
nvcc -shared -o $out/$name.so jaxoplanet_b200/lib/jxp_kernels.o jaxoplanet_b200/lib/jxp_partials.o jaxoplanet_b200/lib/jxp_transit.o jaxoplanet_b200/lib/jxp_state.o $out/$name.w.o -gencode arch=compute_100a,code=sm_100a
rm -f $out/$name.w.o
grep -A2 "k_walk_evalILi.ELb0" $out/$name.log | grep -E "spill|Used" | tr '\n' ' '; echo
