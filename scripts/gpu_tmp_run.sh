mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "batch_vs_oracle or zone_spanning or graph_capturable or step_host_routes" 2>&1 | tail -5
timeout 300 python scripts/overlap_compare.py 16384 65536 262144 1048576 2>&1 | tail -4
echo twostream; IGLU_B200_SO=$PWD/gridworld_b200/csrc/libiglu_b200_twostream.so timeout 300 python scripts/overlap_compare.py 16384 65536 262144 1048576 2>&1 | tail -4
timeout 300 python scripts/tune_r2.py default 2>&1 | tail -1 | cut -c1-220
