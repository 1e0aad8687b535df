mkdir -p gpurun_out
O=gpurun_out/r2d_ab.jsonl
: > $O
run() { echo "# $*" >> $O; timeout 300 "$@" >> $O 2>> gpurun_out/r2d_ab.err; }
timeout 300 python -m pytest tests/test_gpu_solve.py -x -q -k compact > gpurun_out/r2d_pytest.log 2>&1
run python bench.py --skip-extras --steps 10 --in-flight 2
run python bench.py --skip-extras --steps 10 --in-flight 3
run python bench.py --skip-extras --steps 12 --in-flight 4 --shards 4
run env MIQP_NO_WIDE_TEAM=1 python bench.py --skip-extras --steps 10 --in-flight 2
run python bench.py --skip-extras --steps 12 --in-flight 4 --shards 4 --batch 1024
tail -3 gpurun_out/r2d_pytest.log
