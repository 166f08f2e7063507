#!/bin/bash
# Copies the captures of scripts/gpu_final.sh from gpurun_out/ into profiles/ (tracked). usage: collect_profiles.sh r01
r="${1:-r01}"
cp gpurun_out/final_bench_line.json profiles/${r}_bench_line.json
cp gpurun_out/final_launches.csv profiles/${r}_bench_launches.csv
{ echo "ncu --metrics gpu__time_duration.sum --clock-control none -c 400, python bench.py --steps 150 --warmup 10 --no-cpu --no-mi (per-launch times are serialised; shares are what counts)"
  python scripts/launch_summary.py gpurun_out/final_launches.csv; } > profiles/${r}_bench_launches_summary.txt
{ echo "ncu --set full --clock-control none --import-source on -k regex:k_step -s 700 -c 2, python scripts/sweep.py one every_step 0 65536 400  (steady state: step ~350 of the config-2 workload)"
  python scripts/ncu_summary.py gpurun_out/prof_final_steady.ncu-rep; } > profiles/${r}_step_kernels_ncu_summary.txt
python scripts/ncu_lines.py gpurun_out/prof_final_steady.ncu-rep '(int)2>' > profiles/${r}_observe_kernel_ncu_lines.txt 2>&1
python scripts/ncu_lines.py gpurun_out/prof_final_steady.ncu-rep '(int)1>' > profiles/${r}_physics_kernel_ncu_lines.txt 2>&1
cp gpurun_out/final_sweep.jsonl profiles/${r}_sweep_walking_flying_goals.jsonl
cp gpurun_out/final_sweep_reward.jsonl profiles/${r}_reward_sweep.jsonl
