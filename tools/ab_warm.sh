#!/bin/bash
for MU in 0 1 0.1 0.01 0.001; do
  echo "== WARM_MU=$MU"
  MIQP_WARM_MU=$MU timeout 200 python tools/round_trace.py --batch 2048 2>&1 | grep -v "^\[miqp" | head -3
done
