mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2b_gpu.txt
timeout 600 python -m pytest tests/test_gpu_frontier.py -x -q > gpurun_out/r2b_frontier_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2b_frontier_pytest.log
timeout 200 python bench.py --workload config5 --steps 1 --warmup 1 --time-limit 3 --cpu-sample 0 > gpurun_out/r2b_c5.json 2> gpurun_out/r2b_c5.err
timeout 200 python bench.py --workload config5 --cars 4 --horizon 20 --steps 1 --warmup 1 --time-limit 3 --cpu-sample 0 > gpurun_out/r2b_c5_4x20.json 2> gpurun_out/r2b_c5_4x20.err
timeout 300 python bench.py --workload config4 --steps 1 --warmup 1 --batch 256 --cpu-sample 4 > gpurun_out/r2b_c4.json 2> gpurun_out/r2b_c4.err
timeout 300 python bench.py --workload config3 --steps 1 --warmup 1 --batch 64 --cpu-sample 4 > gpurun_out/r2b_c3.json 2> gpurun_out/r2b_c3.err
tail -3 gpurun_out/r2b_frontier_pytest.log
