#!/bin/bash
# ncu --set full of the two step kernels at steady state for the default build (and optionally one variant library).
# usage: gpu_prof_step.sh tag [variant]
tag=$1; var=$2
mkdir -p gpurun_out
run() {  # $1 = output stem
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_step|k_observe" -s 570 -c 2 \
      -f -o gpurun_out/$1 python bench.py --steps 20 --warmup 5 --no-cpu --no-mi --legs noconfig3,nodropin,nodense > gpurun_out/$1.log 2>&1
  python scripts/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1_summary.txt 2>&1
  python scripts/ncu_lines.py gpurun_out/$1.ncu-rep 'k_observe' 70 65536 > gpurun_out/$1_observe_lines.txt 2>&1
  python scripts/ncu_lines.py gpurun_out/$1.ncu-rep '(int)2>' 70 65536 >> gpurun_out/$1_observe_lines.txt 2>&1
  python scripts/ncu_lines.py gpurun_out/$1.ncu-rep '(int)1>' 60 65536 > gpurun_out/$1_physics_lines.txt 2>&1
}
run ${tag}_prof_step
if [ -n "$var" ]; then IGLU_B200_SO=$PWD/gridworld_b200/csrc/libiglu_b200_$var.so run ${tag}_prof_step_$var; fi
grep -E "KERNEL|duration|inst_executed.sum|issue_active|dram__bytes" gpurun_out/${tag}_prof_step*_summary.txt
