#!/bin/bash
# Final captures of a round: bench line, ncu launch list of the same command, steady-state ncu reports, sweeps.
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/final_bench.log 2>&1; tail -1 gpurun_out/final_bench.log > gpurun_out/final_bench_line.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/final_launches.csv python bench.py --steps 150 --warmup 10 --no-cpu --no-mi > gpurun_out/final_ncu_launches.log 2>&1
bash scripts/gpu_prof_steady.sh final_steady every_step 0 65536
timeout 900 python scripts/sweep.py walking flying goals > gpurun_out/final_sweep.jsonl 2>&1
timeout 600 python scripts/sweep.py reward > gpurun_out/final_sweep_reward.jsonl 2>&1
tail -3 gpurun_out/final_sweep.jsonl | cut -c1-200
