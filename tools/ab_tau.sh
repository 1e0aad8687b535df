#!/bin/bash
for TK in 0 0.1 0.3 1 3; do
  echo "== TAU_K=$TK"
  MIQP_TAU_K=$TK timeout 200 python tools/round_trace.py --batch 2048 2>&1 | grep -v "^\[miqp" | sed -n 1,4p
done
