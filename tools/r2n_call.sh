mkdir -p gpurun_out
nproc > gpurun_out/r2n_host.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
( time timeout 900 $TR bench.py --gpus 8 --steps 9 --warmup 3 > gpurun_out/r2n_bench_8gpu.json 2> gpurun_out/r2n_bench_8gpu.err ) 2> gpurun_out/r2n_time.txt
timeout 300 $TR bench.py --gpus 8 --workload config5 --steps 1 --warmup 1 --time-limit 5 --cpu-sample 0 > gpurun_out/r2n_c5_8gpu.json 2> gpurun_out/r2n_c5_8gpu.err
grep -c "NCCL INFO" gpurun_out/r2n_bench_8gpu.err; grep -i "nranks" gpurun_out/r2n_bench_8gpu.err | head -2 | cut -c1-200
cat gpurun_out/r2n_time.txt
