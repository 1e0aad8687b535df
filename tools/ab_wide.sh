#!/bin/bash
for WD in 0 2 4 8; do
  echo "== WIDE_DIV=$WD"
  for B in 2048 1024; do MIQP_WIDE_DIV=$WD timeout 200 python tools/round_trace.py --batch $B 2>&1 | grep -v "^\[miqp" | head -2; done
done
