#!/bin/bash
# usage: gpu_prof_one.sh <name> <reward_mode> <inplace> <envs> [kernel-regex]
mkdir -p gpurun_out
kre="${5:-k_step}"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s 30 -c 1 \
    -f -o gpurun_out/prof_$1 python scripts/sweep.py one $2 $3 $4 40 > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log
