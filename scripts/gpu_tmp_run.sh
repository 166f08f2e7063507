timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "reward" 2>&1 | tail -4
timeout 600 python scripts/mi_full_support.py
timeout 300 python scripts/tune_r2.py default 2>&1 | tail -1 | cut -c150-260
