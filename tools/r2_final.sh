#!/bin/bash
# One GPU-box pass of the round's final state: parity tests, bench lines of every workload, reference arm, ncu launch list.
TAG=${1:-r2z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt
nproc >> gpurun_out/${TAG}_gpu.txt; lscpu | grep "Model name" >> gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 9 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --workload config4 --steps 2 --warmup 1 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err
timeout 300 python bench.py --workload config3 --steps 2 --warmup 1 > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
timeout 300 python bench.py --workload config5 --steps 1 --warmup 1 --time-limit 10 > gpurun_out/${TAG}_c5.json 2> gpurun_out/${TAG}_c5.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 2 --warmup 1 --in-flight 1 --skip-extras > gpurun_out/${TAG}_ncu_bench.log 2>&1
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
