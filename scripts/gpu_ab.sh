#!/bin/bash
# A/B of two builds of the library on the same box: build_ab/head.so (previous commit) vs the in-tree build.
for m in "on_change 2 65536" "on_change 2 1048576" "on_change 0 65536" "every_step 0 65536" "every_step 2 65536"; do
  for rep in 1 2; do
    echo "A(prev) $m: $(IGLU_B200_SO=$PWD/build_ab/head.so python scripts/sweep.py one $m 300 | sed 's/.*"ms_per_step": \([0-9.]*\).*/\1/')"
    echo "B(new)  $m: $(python scripts/sweep.py one $m 300 | sed 's/.*"ms_per_step": \([0-9.]*\).*/\1/')"
  done
done
