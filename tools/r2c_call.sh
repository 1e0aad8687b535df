mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2c_pytest.log
timeout 600 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
timeout 200 python bench.py --workload config5 --steps 1 --warmup 1 --time-limit 5 --cpu-sample 0 > gpurun_out/r2c_c5.json 2> gpurun_out/r2c_c5.err
timeout 200 python bench.py --workload config5 --cars 4 --horizon 20 --steps 1 --warmup 1 --time-limit 3 --cpu-sample 0 > gpurun_out/r2c_c5_4x20.json 2> gpurun_out/r2c_c5_4x20.err
timeout 300 python bench.py --workload config4 --steps 1 --warmup 1 --batch 256 --cpu-sample 4 > gpurun_out/r2c_c4.json 2> gpurun_out/r2c_c4.err
tail -5 gpurun_out/r2c_pytest.log
