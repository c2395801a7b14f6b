mkdir -p gpurun_out
(timeout 200 python scripts/cluster_trace.py long; timeout 200 python scripts/cluster_trace.py group) 2>&1 | tee gpurun_out/cluster_trace.txt
