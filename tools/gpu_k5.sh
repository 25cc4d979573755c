cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "transit_partials" 2>&1 | tail -5
timeout 600 python tools/c5_store_bench.py 2>&1 | tail -12
