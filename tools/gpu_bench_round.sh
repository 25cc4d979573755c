#!/bin/bash
# one gpurun session: GPU tests, bench.py (plain), then the ncu launch list and one --set full capture of the
# dominant kernel of the SAME bench command (each only after the plain run exited 0)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-m}; kern=${2:-k_walk_eval}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_${tag}_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r2_${tag}_tests.log
tail -4 gpurun_out/r2_${tag}_tests.log
timeout 400 python bench.py > gpurun_out/r2_${tag}_bench.json 2> gpurun_out/r2_${tag}_bench.err || { tail -5 gpurun_out/r2_${tag}_bench.err; exit 1; }
cut -c1-300 gpurun_out/r2_${tag}_bench.json
BCMD="python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu"
timeout 300 $BCMD > gpurun_out/r2_${tag}_plain.json 2> gpurun_out/r2_${tag}_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_${tag}_launches.csv $BCMD > gpurun_out/r2_${tag}_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s 12 -c 1 -o gpurun_out/r2_${tag}_${kern} -f $BCMD > gpurun_out/r2_${tag}_ncu_f.log 2>&1
tail -2 gpurun_out/r2_${tag}_ncu_f.log
# K5 with rows stored (config-5 shape, 4096 sets): plain run, then one --set full capture
timeout 300 python tools/k5_prof.py 4096 f64 > gpurun_out/r2_${tag}_k5.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_transit_partials -s 1 -c 1 -o gpurun_out/r2_${tag}_k5 -f python tools/k5_prof.py 4096 f64 >> gpurun_out/r2_${tag}_k5.log 2>&1
tail -3 gpurun_out/r2_${tag}_k5.log
