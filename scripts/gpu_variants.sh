#!/bin/bash
# times the tuning variants built by gridworld_b200.build.build_variant
for v in "$@"; do
  so=$PWD/gridworld_b200/csrc/libiglu_b200_$v.so
  echo "== $v"
  IGLU_B200_SO=$so python scripts/sweep.py one every_step 0 65536 300
  IGLU_B200_SO=$so python scripts/sweep.py one every_step 2 65536 300
  IGLU_B200_SO=$so python scripts/sweep.py one every_step 0 1048576 30
done
