mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2h_gpus.txt
timeout 600 python -m pytest tests/test_gpu_frontier.py -x -q > gpurun_out/r2h_frontier_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2h_frontier_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --workload config5 --steps 1 --warmup 1 --time-limit 5 --cpu-sample 0 > gpurun_out/r2h_c5_2gpu.json 2> gpurun_out/r2h_c5_2gpu.err
timeout 300 $TR bench.py --gpus 2 --workload config5 --cars 4 --horizon 20 --steps 1 --warmup 1 --time-limit 3 --cpu-sample 0 > gpurun_out/r2h_c5_4x20_2gpu.json 2> gpurun_out/r2h_c5_4x20_2gpu.err
timeout 400 $TR bench.py --gpus 2 --skip-extras --steps 9 --in-flight 3 > gpurun_out/r2h_bench_2gpu.json 2> gpurun_out/r2h_bench_2gpu.err
grep -c "NCCL INFO" gpurun_out/r2h_c5_2gpu.err; grep -i "nranks\|Init COMPLETE" gpurun_out/r2h_c5_2gpu.err | head -4
tail -3 gpurun_out/r2h_frontier_pytest.log
