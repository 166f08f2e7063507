mkdir -p gpurun_out; : > gpurun_out/r2u_tune.jsonl
timeout 300 python scripts/tune_r2.py default >> gpurun_out/r2u_tune.jsonl 2>gpurun_out/r2u_tune.err
for v in p5w16 p6w16 p5w12 p6w12 p8w16; do
  IGLU_B200_SO=$PWD/gridworld_b200/csrc/libiglu_b200_$v.so timeout 300 python scripts/tune_r2.py $v >> gpurun_out/r2u_tune.jsonl 2>>gpurun_out/r2u_tune.err
done
cut -c1-200 gpurun_out/r2u_tune.jsonl; tail -3 gpurun_out/r2u_tune.err
