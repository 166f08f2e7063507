#!/bin/bash
# Round-2 evidence: launch list of the bench command, ncu --set full of the three step kernels at steady state and of
# the stateless sweep, per-line summaries, SASS excerpt of the TMA / mbarrier instructions. Results land in gpurun_out/
# (copied to profiles/ by scripts/collect_profiles.sh once read).
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-mi --legs noconfig3,nodropin,nopov"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv \
    --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
# steady state: the pre-roll is 280 steps x 3 step kernels; take the 3 kernels of one step after it
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_step|k_widen|k_commit" -s 870 -c 3 \
    -f -o gpurun_out/r02_prof_step $B > gpurun_out/r02_ncu_step.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_prof_step.ncu-rep > gpurun_out/r02_step_kernels_ncu_summary.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r02_prof_step.ncu-rep 'k_widen' 70 65536 > gpurun_out/r02_widen_kernel_ncu_lines.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r02_prof_step.ncu-rep '(int)1>' 60 65536 > gpurun_out/r02_physics_kernel_ncu_lines.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r02_prof_step.ncu-rep 'k_commit' 40 4600 > gpurun_out/r02_commit_kernel_ncu_lines.txt 2>&1
for w in "compact 0.2" "goals 0.2"; do
  set -- $w
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_max_intersection" -s 1 -c 1 \
      -f -o gpurun_out/r02_prof_mi_$1 python scripts/prof_reward.py $1 $2 > gpurun_out/r02_ncu_mi_$1.log 2>&1
  python scripts/ncu_summary.py gpurun_out/r02_prof_mi_$1.ncu-rep > gpurun_out/r02_max_intersection_$1_ncu_summary.txt 2>&1
  python scripts/ncu_lines.py gpurun_out/r02_prof_mi_$1.ncu-rep 'k_max_intersection' 60 1000000 > gpurun_out/r02_max_intersection_$1_ncu_lines.txt 2>&1
done
cuobjdump -sass gridworld_b200/csrc/libiglu_b200.so | grep -E "Function|UBLKCP|SYNCS|REDS|ATOMS" | grep -B1 -E "UBLKCP|SYNCS" | head -80 > gpurun_out/r02_sass_tma.txt
timeout 300 python scripts/overlap_trace.py 65536 > gpurun_out/r02_overlap_trace.jsonl 2>&1
tail -4 gpurun_out/r02_launches_summary.txt; grep -E "KERNEL|duration|dram__bytes|inst_executed.sum|issue_active" gpurun_out/r02_step_kernels_ncu_summary.txt
