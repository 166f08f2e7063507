#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 \
    -f -o gpurun_out/prof_phys python scripts/sweep.py one on_change 1 65536 30 > gpurun_out/ncu_phys.log 2>&1
tail -2 gpurun_out/ncu_phys.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 \
    -f -o gpurun_out/prof_obs python scripts/sweep.py one on_change 0 65536 30 > gpurun_out/ncu_obs.log 2>&1
tail -2 gpurun_out/ncu_obs.log
