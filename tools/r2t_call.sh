mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555"
timeout 600 $TR bench.py --gpus 8 --steps 9 --warmup 3 > gpurun_out/r2t_bench_8gpu.json 2> gpurun_out/r2t_bench_8gpu.err
tail -c 300 gpurun_out/r2t_bench_8gpu.json
