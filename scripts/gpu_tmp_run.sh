mkdir -p gpurun_out
timeout 600 python scripts/obs_compare.py 16384 65536 262144 1048576 > gpurun_out/r2l_obs_compare.jsonl 2> gpurun_out/r2l_obs_compare.err
cat gpurun_out/r2l_obs_compare.jsonl; tail -3 gpurun_out/r2l_obs_compare.err
