#!/bin/bash
# build a named variant of the library with extra nvcc flags: tools/build_variant.sh NAME "-DX=1 ..."
set -e
cd "$(dirname "$0")/.."
out=jaxoplanet_b200/lib/variants; mkdir -p $out
name=$1; shift
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --fmad=true -Xptxas -v $*"
nvcc $F -c jaxoplanet_b200/csrc/jxp_kernels.cu -o $out/$name.k.o > $out/$name.log 2>&1 &
nvcc $F -c jaxoplanet_b200/csrc/jxp_partials.cu -o $out/$name.p.o > $out/$name.p.log 2>&1 &
nvcc $F -c jaxoplanet_b200/csrc/jxp_transit.cu -o $out/$name.t.o > $out/$name.t.log 2>&1 &
wait
nvcc -shared -o $out/$name.so $out/$name.k.o $out/$name.p.o $out/$name.t.o -gencode arch=compute_100a,code=sm_100a
rm -f $out/$name.k.o $out/$name.p.o $out/$name.t.o
grep -A3 "k_systemIdLi11ELb1ELb0ELb0" $out/$name.log | grep -E "spill|Used"
