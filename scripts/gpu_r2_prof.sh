#!/bin/bash
# round 2 profiling call: launch list of the bench command, ncu --set full of the step kernels at steady state and of the
# stateless sweep, summaries into gpurun_out/ (copied to profiles/ by hand once read).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 400 --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-mi --legs noconfig3,nodropin > gpurun_out/r2_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_summary.txt 2>&1
# steady state: skip the pre-roll (250 masked resets + 280 steps x 2 launches = ~810 launches) and the set-up
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_step" -s 900 -c 2 \
    -f -o gpurun_out/r2_prof_step python bench.py --steps 20 --warmup 5 --no-cpu --no-mi --legs noconfig3,nodropin > gpurun_out/r2_ncu_step.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_prof_step.ncu-rep > gpurun_out/r2_step_kernels_ncu_summary.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r2_prof_step.ncu-rep '(int)2>' 60 65536 > gpurun_out/r2_observe_kernel_ncu_lines.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r2_prof_step.ncu-rep '(int)1>' 60 65536 > gpurun_out/r2_physics_kernel_ncu_lines.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_max_intersection" -s 1 -c 1 \
    -f -o gpurun_out/r2_prof_mi python scripts/prof_reward.py compact 0.2 > gpurun_out/r2_ncu_mi.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_prof_mi.ncu-rep > gpurun_out/r2_max_intersection_ncu_summary.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r2_prof_mi.ncu-rep 'k_max_intersection' 60 1000000 > gpurun_out/r2_max_intersection_ncu_lines.txt 2>&1
cuobjdump -sass gridworld_b200/csrc/libiglu_b200.so | grep -E "Function|UBLKCP|SYNCS|ATOMS|REDS" | grep -B1 -E "UBLKCP|SYNCS" | head -60 > gpurun_out/r2_sass_tma.txt
tail -3 gpurun_out/r2_launches_summary.txt; head -40 gpurun_out/r2_step_kernels_ncu_summary.txt
