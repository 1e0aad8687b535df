#!/bin/bash
for CFG in "14 8 2" "10 8 2" "14 8 1" "10 16 2" "8 8 2" "12 4 2"; do
  set -- $CFG
  echo "== PATIENCE=$1 GROWTH=$2 DIVE_FILL=$3"
  for B in 2048 1024; do MIQP_DIVE_PATIENCE=$1 MIQP_DIVE_GROWTH=$2 MIQP_DIVE_FILL=$3 timeout 200 python tools/round_trace.py --batch $B 2>&1 | grep -v "^\[miqp" | head -2; done
done
