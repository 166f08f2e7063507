#!/bin/bash
# round 2, second GPU call: parity tests, the bench (new kernels vs the round-1 kernels, A/B), group-size sweep, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60 > gpurun_out/r2_pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --legs nopyref,config5 > gpurun_out/r2_bench_default.log 2>&1
echo "bench exit: $?" >> gpurun_out/r2_bench_default.log
IGLU_B200_SO=$PWD/gridworld_b200/csrc/libiglu_b200_r2base.so timeout 600 python bench.py --legs nopyref,nodropin,config5 --no-cpu > gpurun_out/r2_bench_default_base.log 2>&1
echo "bench exit: $?" >> gpurun_out/r2_bench_default_base.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_driver.log 2>&1
echo "bench exit: $?" >> gpurun_out/r2_bench_driver.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.log 2>&1
G=2,4,8,16 timeout 600 python scripts/sched_sweep.py > gpurun_out/r2_sched_sweep.jsonl 2>&1
bash scripts/gpu_r2_prof.sh > gpurun_out/r2_prof.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu.log; tail -c 600 gpurun_out/r2_bench_default.log; cat gpurun_out/r2_sched_sweep.jsonl | tail -12
