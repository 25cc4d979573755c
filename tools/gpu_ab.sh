#!/bin/bash
cd "$(dirname "$0")/.."
for rep in 1 2; do
echo "== default (working tree)"; python tools/ab_step.py 2>&1 | tail -2
for so in jaxoplanet_b200/lib/variants/*.so; do
  echo "== $(basename $so .so)"; JXP_LIB_PATH=$PWD/$so python tools/ab_step.py 2>&1 | tail -2
done
done
echo "== c4 default"; python tools/c4_bench.py 2>&1 | grep "^walk ms\|^walk per"
for so in jaxoplanet_b200/lib/variants/*.so; do
  echo "== c4 $(basename $so .so)"; JXP_LIB_PATH=$PWD/$so python tools/c4_bench.py 2>&1 | grep "^walk ms\|^walk per"
done
