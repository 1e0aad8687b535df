mkdir -p gpurun_out
nproc > gpurun_out/r2l_host.txt; free -g | head -2 >> gpurun_out/r2l_host.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533"
( time timeout 900 $TR bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/r2l_bench_4gpu.json 2> gpurun_out/r2l_bench_4gpu.err ) 2> gpurun_out/r2l_time.txt
timeout 300 $TR bench.py --gpus 4 --workload config5 --steps 1 --warmup 1 --time-limit 5 --cpu-sample 0 > gpurun_out/r2l_c5_4gpu.json 2> gpurun_out/r2l_c5_4gpu.err
timeout 300 $TR bench.py --gpus 4 --impl reference --steps 2 --warmup 1 > gpurun_out/r2l_ref_4gpu.json 2> gpurun_out/r2l_ref_4gpu.err
grep -c "NCCL INFO" gpurun_out/r2l_bench_4gpu.err; grep -i "nranks" gpurun_out/r2l_bench_4gpu.err | head -3
cat gpurun_out/r2l_time.txt
