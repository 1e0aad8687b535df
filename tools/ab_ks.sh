#!/bin/bash
for KS in 64 128 296; do for DF in 1 2; do
  echo "== KS=$KS DIVE_FILL=$DF"
  MIQP_KS=$KS MIQP_DIVE_FILL=$DF timeout 200 python tools/round_trace.py --batch 2048 2>&1 | grep -v "^\[miqp" | head -2
  MIQP_KS=$KS MIQP_DIVE_FILL=$DF timeout 200 python tools/round_trace.py --batch 1024 2>&1 | grep -v "^\[miqp" | head -2
done; done
