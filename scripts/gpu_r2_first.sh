#!/bin/bash
# round 2, first GPU call: parity tests + smoke + the restructured bench, new kernels vs the round-1 kernels (A/B)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60 > gpurun_out/r2_pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --legs config5 > gpurun_out/r2_bench_driver.log 2>&1
echo "bench exit: $?" >> gpurun_out/r2_bench_driver.log
timeout 600 python bench.py --legs nopyref,config5 > gpurun_out/r2_bench_default.log 2>&1
echo "bench exit: $?" >> gpurun_out/r2_bench_default.log
IGLU_B200_SO=$PWD/gridworld_b200/csrc/libiglu_b200_r2base.so timeout 600 python bench.py --legs nopyref,nodropin,config5 --no-cpu > gpurun_out/r2_bench_default_base.log 2>&1
echo "bench exit: $?" >> gpurun_out/r2_bench_default_base.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.log 2>&1
tail -8 gpurun_out/r2_pytest_gpu.log; tail -2 gpurun_out/r2_smoke.log; tail -c 1200 gpurun_out/r2_bench_default.log
