#!/bin/bash
# A/B on the GPU: parity tests of the current build, then steady-state timings of it and of the variant libraries
# named on the command line (gridworld_b200/csrc/libiglu_b200_<name>.so). usage: scripts/gpu_r2_ab.sh tag [variants...]
tag=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15 > gpurun_out/${tag}_pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/${tag}_pytest_gpu.log
tail -5 gpurun_out/${tag}_pytest_gpu.log
: > gpurun_out/${tag}_tune.jsonl
timeout 300 python scripts/tune_r2.py default >> gpurun_out/${tag}_tune.jsonl 2>gpurun_out/${tag}_tune.err
for v in "$@"; do
  IGLU_B200_SO=$PWD/gridworld_b200/csrc/libiglu_b200_$v.so timeout 300 python scripts/tune_r2.py $v >> gpurun_out/${tag}_tune.jsonl 2>>gpurun_out/${tag}_tune.err
done
cat gpurun_out/${tag}_tune.jsonl; tail -5 gpurun_out/${tag}_tune.err
