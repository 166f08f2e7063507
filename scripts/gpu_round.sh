#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, smoke, bench, sweep, ncu. Logs go to gpurun_out/.
# usage: gpu_round.sh [tests] [bench] [sweep] [ncu]   (default: all)
what="${*:-tests bench sweep ncu}"
mkdir -p gpurun_out
if [[ $what == *tests* ]]; then
  timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
  echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "smoke exit: $?" >> gpurun_out/smoke.log
  tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
fi
if [[ $what == *bench* ]]; then
  timeout 600 python bench.py > gpurun_out/bench.log 2>&1
  echo "bench exit: $?" >> gpurun_out/bench.log
  tail -2 gpurun_out/bench.log
fi
if [[ $what == *sweep* ]]; then
  timeout 900 python scripts/sweep.py > gpurun_out/sweep.log 2>&1
  cat gpurun_out/sweep.log
fi
if [[ $what == *ncu* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-mi > gpurun_out/ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 10 -c 2 \
      -f -o gpurun_out/prof_step python bench.py --steps 10 --warmup 5 --no-cpu --no-mi > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
  ls -la gpurun_out
fi
