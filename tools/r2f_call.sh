mkdir -p gpurun_out
python tools/round_times.py 2> gpurun_out/r2f_rounds_default.log
MIQP_NO_NARROW_TEAM=1 python tools/round_times.py 2> gpurun_out/r2f_rounds_nonarrow.log
timeout 300 python -m pytest tests/test_gpu_solve.py tests/test_gpu_full_batch_parity.py -x -q > gpurun_out/r2f_pytest.log 2>&1
O=gpurun_out/r2f_ab.jsonl
: > $O
run() { echo "# $*" >> $O; timeout 300 "$@" >> $O 2>> gpurun_out/r2f_ab.err; }
run python bench.py --skip-extras --steps 9 --in-flight 3
run env MIQP_NO_NARROW_TEAM=1 python bench.py --skip-extras --steps 9 --in-flight 3
run python bench.py --skip-extras --steps 9 --in-flight 1
tail -3 gpurun_out/r2f_pytest.log
grep "round [1-6]:" gpurun_out/r2f_rounds_default.log gpurun_out/r2f_rounds_nonarrow.log
