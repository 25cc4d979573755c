#!/bin/bash
cd "$(dirname "$0")/.."
f() { python tools/tl_bench.py 2>&1 | python -c "
import json,sys
d=json.load(sys.stdin)
l=d['list']
print({k:(round(v['kernel_ms'],4) if isinstance(v,dict) else v) for k,v in d.items()}, 'step', round(l['step_ms'],4), 'inside', l['inside_items'], 'fast', l['fast_chunks'], 'of', l['inside_items']//64)"; }
echo "== base"; f
for so in jaxoplanet_b200/lib/variants/tl_*.so; do
  echo "== $(basename $so .so)"; JXP_LIB_PATH=$PWD/$so f
done
