mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -6
timeout 300 python scripts/overlap_compare.py 65536 1048576 2>&1 | tail -2
timeout 300 python scripts/tune_r2.py default 2>&1 | tail -1 | cut -c1-220
