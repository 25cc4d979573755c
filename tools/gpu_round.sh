#!/bin/bash
# one gpurun session: tests, walk bench on the base build and on every variant under lib/variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_${tag}_smi.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_${tag}_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r2_${tag}_tests.log
tail -5 gpurun_out/r2_${tag}_tests.log
( echo "== base"; timeout 300 python tools/wk_bench.py
  for so in jaxoplanet_b200/lib/variants/wk_*.so; do
    [ -e "$so" ] || continue
    echo "== $(basename $so .so)"; JXP_LIB_PATH=$PWD/$so timeout 300 python tools/wk_bench.py
  done ) > gpurun_out/r2_${tag}_wk.log 2>&1
cat gpurun_out/r2_${tag}_wk.log | cut -c1-600
timeout 600 python tools/tl_fuzz.py 60 > gpurun_out/r2_${tag}_fuzz.log 2>&1; tail -2 gpurun_out/r2_${tag}_fuzz.log
