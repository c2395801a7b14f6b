mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_cluster -s 3 -c 1 -f -o gpurun_out/r02_cluster_long python scripts/cluster_trace.py long 2>&1 | tail -5
ls -la gpurun_out/*.ncu-rep
