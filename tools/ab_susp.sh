#!/bin/bash
for SB in 0 6 8 10 12 16; do
  echo "== SUSP_BUDGET=$SB"
  for B in 2048 1024; do MIQP_SUSP_BUDGET=$SB timeout 200 python tools/round_trace.py --batch $B 2>&1 | grep -v "^\[miqp" | head -2; done
done
