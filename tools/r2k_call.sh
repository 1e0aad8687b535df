mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pybind_miqp.py tests/test_gpu_configs.py tests/test_gpu_multi.py tests/test_highs_brackets.py tests/test_gpu_frontier.py -x -q > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2k_pytest.log
O=gpurun_out/r2k_ab.jsonl
: > $O
run() { echo "# $*" >> $O; timeout 400 "$@" >> $O 2>> gpurun_out/r2k_ab.err; }
run python bench.py --workload config4 --steps 1 --warmup 1 --batch 512 --cpu-sample 0
run env MIQP_MULTI_PLUNGE=0 python bench.py --workload config4 --steps 1 --warmup 1 --batch 512 --cpu-sample 0
run python bench.py --workload config3 --steps 1 --warmup 1 --batch 256 --cpu-sample 0
run env MIQP_MULTI_PLUNGE=0 python bench.py --workload config3 --steps 1 --warmup 1 --batch 256 --cpu-sample 0
run python bench.py --skip-extras --steps 9 --in-flight 3 --batch 4096
run python bench.py --skip-extras --steps 12 --in-flight 4 --shards 4
tail -5 gpurun_out/r2k_pytest.log
