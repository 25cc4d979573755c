#!/bin/bash
cd "$(dirname "$0")/.."
echo "== base"; python tools/tl_bench.py 2>&1 | grep -A3 '"list"' | tr -d '\n'; echo
for so in jaxoplanet_b200/lib/variants/tl_*.so; do
  echo "== $(basename $so .so)"
  JXP_LIB_PATH=$PWD/$so python tools/tl_bench.py 2>&1 | grep -A3 '"list"' | tr -d '\n'; echo
done
