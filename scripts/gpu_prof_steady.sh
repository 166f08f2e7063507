#!/bin/bash
# steady-state ncu capture of both step kernels (k_step<*,1> physics, k_step<*,2> observe) after ~350 steps
# usage: gpu_prof_steady.sh <name> <reward_mode> <inplace> <envs>
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_step" -s 700 -c 2 \
    -f -o gpurun_out/prof_$1 python scripts/sweep.py one $2 $3 $4 400 > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log
