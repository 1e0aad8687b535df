#!/bin/bash
for PAT in 0 10 14 18; do for GR in 2 8; do
  echo "== PATIENCE=$PAT GROWTH=$GR"
  for B in 2048 1024; do MIQP_DIVE_PATIENCE=$PAT MIQP_DIVE_GROWTH=$GR timeout 200 python tools/round_trace.py --batch $B 2>&1 | grep -v "^\[miqp" | head -2; done
done; done
