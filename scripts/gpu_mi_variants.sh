#!/bin/bash
# stateless reward sweep with different warps-per-CTA x CTAs-per-SM of k_max_intersection
for v in mi8x2 mi10x2 default mi24x1 mi27x1; do
  so=""; [ $v != default ] && so="$PWD/gridworld_b200/csrc/libiglu_b200_$v.so"
  echo "== $v"
  IGLU_B200_SO=$so python scripts/sweep.py reward 2>&1 | sed 's/.*"support": "\([a-z]*\)", "rho": \([0-9.]*\), "ms": \([0-9.]*\).*/\1 \2 \3/'
done
