#!/bin/bash
# ncu --set full capture of one k_walk_eval launch inside tools/wk_bench.py (NS=4096), after the plain run exited 0
cd "$(dirname "$0")/.."
tag=${1:-a}; kern=${2:-k_walk_eval}; lib=${3:-}
[ -n "$lib" ] && export JXP_LIB_PATH=$PWD/$lib
NS=4096 timeout 300 python tools/wk_bench.py > gpurun_out/r2_${tag}_plain.log 2>&1 || exit 1
NS=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s 6 -c 1 \
  -o gpurun_out/r2_${tag} -f python tools/wk_bench.py > gpurun_out/r2_${tag}_ncu.log 2>&1
tail -3 gpurun_out/r2_${tag}_ncu.log
