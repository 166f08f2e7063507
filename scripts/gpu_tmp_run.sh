mkdir -p gpurun_out; : > gpurun_out/r2w_tune.jsonl
timeout 300 python scripts/tune_r2.py default >> gpurun_out/r2w_tune.jsonl 2>gpurun_out/r2w_tune.err
for v in w56p5 w48p5 w48p6 w40p5; do
  IGLU_B200_SO=$PWD/gridworld_b200/csrc/libiglu_b200_$v.so timeout 300 python scripts/tune_r2.py $v >> gpurun_out/r2w_tune.jsonl 2>>gpurun_out/r2w_tune.err
done
cut -c1-200 gpurun_out/r2w_tune.jsonl; tail -3 gpurun_out/r2w_tune.err
for v in w56p5 w48p5; do echo $v; IGLU_B200_SO=$PWD/gridworld_b200/csrc/libiglu_b200_$v.so timeout 300 python scripts/overlap_compare.py 262144 1048576 2>&1 | tail -2 | cut -c1-120; done
