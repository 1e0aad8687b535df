#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, one full capture of the node kernel.
# usage: tools/gpu_check.sh <tag>
TAG=${1:-rX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt
nproc >> gpurun_out/${TAG}_gpu.txt; lscpu | grep "Model name" >> gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 1 --warmup 1 --batch 2048 --cpu-sample 2 --latency-plans 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bnb_nodes -s 6 -c 1 -o gpurun_out/${TAG}_nodes -f \
   python tools/profile_run.py --batch 2048 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_nodes.ncu-rep --page raw --csv > gpurun_out/${TAG}_nodes_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
