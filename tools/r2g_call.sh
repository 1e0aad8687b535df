mkdir -p gpurun_out
O=gpurun_out/r2g_ab.jsonl
: > $O
run() { echo "# $*" >> $O; timeout 300 "$@" >> $O 2>> gpurun_out/r2g_ab.err; }
B="python bench.py --skip-extras --steps 9 --in-flight 3 --first-shard 100"
run $B
run env MIQP_NARROW_MIN=370 $B
run env MIQP_SUSP_BUDGET=12 $B
run env MIQP_SUSP_BUDGET=16 $B
run env MIQP_SUSP_BUDGET=6 $B
run env MIQP_DIVE_FILL=4 $B
run env MIQP_DIVE_FILL=1 $B
run env MIQP_KS=32 $B
run env MIQP_DIVE_PATIENCE=20 $B
run env MIQP_NO_WIDE_TEAM=1 $B
