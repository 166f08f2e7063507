#!/bin/bash
# round 2, third GPU call: parity tests on the compaction-based evaluation, occupancy / group-size variants, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2c_pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2c_pytest_gpu.log
tail -4 gpurun_out/r2c_pytest_gpu.log
: > gpurun_out/r2c_tune.jsonl
timeout 300 python scripts/tune_r2.py default >> gpurun_out/r2c_tune.jsonl 2>gpurun_out/r2c_tune.err
for g in 2 4 8; do timeout 300 python scripts/tune_r2.py g$g $((g*256)) >> gpurun_out/r2c_tune.jsonl 2>>gpurun_out/r2c_tune.err; done
for v in r2base w24c1r2 w24c1r3 w12c2r3 w16c2r3 w8c4r2 w20c1r4; do
  IGLU_B200_SO=$PWD/gridworld_b200/csrc/libiglu_b200_$v.so timeout 300 python scripts/tune_r2.py $v >> gpurun_out/r2c_tune.jsonl 2>>gpurun_out/r2c_tune.err
done
cat gpurun_out/r2c_tune.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_step" -s 570 -c 2 \
    -f -o gpurun_out/r2c_prof_step python bench.py --steps 20 --warmup 5 --no-cpu --no-mi --legs noconfig3,nodropin > gpurun_out/r2c_ncu_step.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2c_prof_step.ncu-rep > gpurun_out/r2c_step_kernels_ncu_summary.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r2c_prof_step.ncu-rep '(int)2>' 70 65536 > gpurun_out/r2c_observe_kernel_ncu_lines.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r2c_prof_step.ncu-rep '(int)1>' 60 65536 > gpurun_out/r2c_physics_kernel_ncu_lines.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_max_intersection" -s 1 -c 1 \
    -f -o gpurun_out/r2c_prof_mi python scripts/prof_reward.py compact 0.2 > gpurun_out/r2c_ncu_mi.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2c_prof_mi.ncu-rep > gpurun_out/r2c_max_intersection_ncu_summary.txt 2>&1
python scripts/ncu_lines.py gpurun_out/r2c_prof_mi.ncu-rep 'k_max_intersection' 70 200000 > gpurun_out/r2c_max_intersection_ncu_lines.txt 2>&1
head -30 gpurun_out/r2c_step_kernels_ncu_summary.txt
