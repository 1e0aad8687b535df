mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bnb_nodes -s 4 -c 1 -o gpurun_out/r2i_nodes -f \
   python tools/profile_run.py --batch 2048 > gpurun_out/r2i_ncu_full.log 2>&1
ncu -i gpurun_out/r2i_nodes.ncu-rep --page raw --csv > gpurun_out/r2i_nodes_raw.csv 2>/dev/null
ncu -i gpurun_out/r2i_nodes.ncu-rep --page source --csv > gpurun_out/r2i_nodes_source.csv 2>/dev/null
ls -la gpurun_out/r2i*
tail -3 gpurun_out/r2i_ncu_full.log
