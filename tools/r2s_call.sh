mkdir -p gpurun_out
O=gpurun_out/r2s_ab.jsonl
: > $O
run() { echo "# $*" >> $O; timeout 300 "$@" >> $O 2>> gpurun_out/r2s_ab.err; }
B="python bench.py --skip-extras --steps 6 --in-flight 3 --first-shard 100"
run $B
run env MIQP_WARM_MU=1 $B
run env MIQP_WARM_MU=3 $B
run env MIQP_WARM_MU=30 $B
