#!/bin/bash
# tab_check timing (config 3, NS sets) of the base library and of every walk variant under lib/variants
cd "$(dirname "$0")/.."
tag=${1:-x}
( echo "== base"; NS=${NS:-4096,512} NOADV=1 timeout 300 python tools/tab_check.py
  for so in jaxoplanet_b200/lib/variants/wk_*.so; do
    [ -e "$so" ] || continue
    echo "== $(basename $so .so)"; NS=${NS:-4096,512} NOADV=1 JXP_LIB_PATH=$PWD/$so timeout 300 python tools/tab_check.py
  done ) > gpurun_out/r2_${tag}_tabvar.log 2>&1
grep -E "^==|n_sets" gpurun_out/r2_${tag}_tabvar.log | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('=='): print(ln.strip()); continue
    d=json.loads(ln); t=d['tab']; print('   ns',d['n_sets'],'step',t['step_ms'],'items',t['items_ms'],'plan',t['plan_ms'],'tables',t['tables_ms'],'eval',t['eval_ms'],'| direct step',d['direct']['step_ms'],'lldiff',d['ll_abs_diff_tab_vs_direct'])
"
