#!/bin/bash
cd "$(dirname "$0")/.."
for so in jaxoplanet_b200/lib/variants/*.so; do
  echo "== $(basename $so .so)"
  JXP_LIB_PATH=$PWD/$so python tools/prof_case.py c3 4096 100000 0 5 2>&1 | tail -1
done
