#!/bin/bash
# compile jxp_walk.cu alone (fast iteration): tools/cc_walk.sh [extra nvcc flags]
cd /root/repo/jaxoplanet_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --fmad=true -Xptxas -v "$@" -c jxp_walk.cu -o /tmp/jxp_walk.o 2>&1 | grep -A2 "k_walk_eval\|error" | grep -v "^--"
