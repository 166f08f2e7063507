#!/bin/bash
# steady-state timings (scripts/tune_r2.py) of the default build and of the named variant libraries. usage: gpu_tune_only.sh tag [variants...]
tag=$1; shift
mkdir -p gpurun_out
: > gpurun_out/${tag}_tune.jsonl
timeout 300 python scripts/tune_r2.py default >> gpurun_out/${tag}_tune.jsonl 2>gpurun_out/${tag}_tune.err
for v in "$@"; do
  IGLU_B200_SO=$PWD/gridworld_b200/csrc/libiglu_b200_$v.so timeout 300 python scripts/tune_r2.py $v >> gpurun_out/${tag}_tune.jsonl 2>>gpurun_out/${tag}_tune.err
done
cat gpurun_out/${tag}_tune.jsonl; tail -5 gpurun_out/${tag}_tune.err
