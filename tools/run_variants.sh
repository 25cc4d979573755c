#!/bin/bash
# time prof_case c3 / c4 for every library variant under jaxoplanet_b200/lib/variants
cd "$(dirname "$0")/.."
for so in jaxoplanet_b200/lib/variants/*.so; do
  n=$(basename $so .so)
  echo "== $n"
  JXP_LIB_PATH=$PWD/$so python tools/prof_case.py c3 4096 100000 0 5 2>&1 | tail -1
  JXP_LIB_PATH=$PWD/$so python tools/prof_case.py c4 1024 200000 0 3 2>&1 | tail -1
done
