mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -8 > gpurun_out/r2o_pytest.log; cat gpurun_out/r2o_pytest.log
timeout 600 python scripts/overlap_compare.py > gpurun_out/r2o_overlap_compare.jsonl 2> gpurun_out/r2o_overlap_compare.err
cat gpurun_out/r2o_overlap_compare.jsonl; tail -3 gpurun_out/r2o_overlap_compare.err
timeout 300 python scripts/tune_r2.py default > gpurun_out/r2o_tune.jsonl 2>gpurun_out/r2o_tune.err; cut -c1-220 gpurun_out/r2o_tune.jsonl
