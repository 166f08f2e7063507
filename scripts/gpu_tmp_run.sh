mkdir -p gpurun_out; : > gpurun_out/r2x_tune.jsonl
timeout 300 python scripts/tune_r2.py default >> gpurun_out/r2x_tune.jsonl 2>gpurun_out/r2x_tune.err
for v in pw2_w12r56 pw2_w14r48 pw2_w16r40 pw2_w16r56; do
  IGLU_B200_SO=$PWD/gridworld_b200/csrc/libiglu_b200_$v.so timeout 300 python scripts/tune_r2.py $v >> gpurun_out/r2x_tune.jsonl 2>>gpurun_out/r2x_tune.err
done
cut -c1-200 gpurun_out/r2x_tune.jsonl; tail -3 gpurun_out/r2x_tune.err
