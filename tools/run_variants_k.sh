#!/bin/bash
cd "$(dirname "$0")/.."
for so in jaxoplanet_b200/lib/variants/*.so; do
  n=$(basename $so .so)
  echo "== $n"
  JXP_LIB_PATH=$PWD/$so python tools/prof_case.py kepler 1000 100000 0 5 2>&1 | tail -1
done
