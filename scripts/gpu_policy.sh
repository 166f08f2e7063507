#!/bin/bash
# fused vs two-launch step for on_change reward (obs 0 = full rewrite, 2 = incremental) at three sizes
for n in 65536 262144 1048576; do
  for obs in 0 2; do
    echo "fused  on_change obs=$obs n=$n: $(IGLU_NO_SPLIT=1 python scripts/sweep.py one on_change $obs $n 300 | tail -1 | cut -c1-300)"
    echo "split  on_change obs=$obs n=$n: $(IGLU_FORCE_SPLIT=1 python scripts/sweep.py one on_change $obs $n 300 | tail -1 | cut -c1-300)"
  done
done
