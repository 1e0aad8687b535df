mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2e_pytest.log
O=gpurun_out/r2e_ab.jsonl
: > $O
run() { echo "# $*" >> $O; timeout 300 "$@" >> $O 2>> gpurun_out/r2e_ab.err; }
run python bench.py --skip-extras --steps 9 --in-flight 3
run env MIQP_NO_NARROW_TEAM=1 python bench.py --skip-extras --steps 9 --in-flight 3
run env MIQP_NARROW_MIN=600 python bench.py --skip-extras --steps 9 --in-flight 3
run env MIQP_B200_VARIANT=t3 MIQP_NO_NARROW_TEAM=1 python bench.py --skip-extras --steps 9 --in-flight 3
run env MIQP_B200_VARIANT=t3 python bench.py --skip-extras --steps 9 --in-flight 3
tail -3 gpurun_out/r2e_pytest.log
