#!/bin/bash
cd "$(dirname "$0")/.."
echo "== default"; python tools/k5_split.py 2>&1 | tail -3
for so in jaxoplanet_b200/lib/variants/*.so; do
  echo "== $(basename $so .so)"
  JXP_LIB_PATH=$PWD/$so python tools/k5_split.py 2>&1 | tail -3
done
