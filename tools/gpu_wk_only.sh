#!/bin/bash
cd "$(dirname "$0")/.."
tag=${1:-x}
( echo "== base"; NS=${NS:-4096} timeout 300 python tools/wk_bench.py
  for so in jaxoplanet_b200/lib/variants/wk_*.so; do
    [ -e "$so" ] || continue
    echo "== $(basename $so .so)"; NS=${NS:-4096} JXP_LIB_PATH=$PWD/$so timeout 300 python tools/wk_bench.py
  done ) > gpurun_out/r2_${tag}_wk.log 2>&1
python - <<'PY'
import json,sys
name=None
for ln in open('gpurun_out/r2_%s_wk.log' % sys.argv[1] if len(sys.argv)>1 else 'x'):
    pass
PY
grep -E "^==|n_sets" gpurun_out/r2_${tag}_wk.log | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('=='): print(ln.strip(), end=' '); continue
    d=json.loads(ln); print('ns',d['n_sets'],'eval',d['walk']['kernel_ms'],'plan',d['walk_plan']['kernel_ms'],'step',d['walk']['step_ms'],'b2b',d['walk']['b2b_ms'],'window',d['window']['step_ms'],'diff',d['max_rel_diff_vs_window'])
"
