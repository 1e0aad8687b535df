#!/bin/bash
for FM in 1 2 3; do for DF in 2 4; do
  echo "== FILL_MULT=$FM DIVE_FILL=$DF"
  MIQP_FILL_MULT=$FM MIQP_DIVE_FILL=$DF timeout 200 python tools/round_trace.py --batch 2048 2>/dev/null | head -2
done; done
