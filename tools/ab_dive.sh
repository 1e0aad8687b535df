#!/bin/bash
for F in 0 1 2 4; do
  echo "== MIQP_DIVE_FILL=$F"
  MIQP_DIVE_FILL=$F timeout 200 python tools/round_trace.py --batch 2048 2>gpurun_out/trace_$F.log | tail -3
  grep -c "round" gpurun_out/trace_$F.log
done
echo "== batch 8192, fill 0"
MIQP_DIVE_FILL=0 timeout 300 python tools/round_trace.py --batch 8192 2>/dev/null | tail -3
