mkdir -p gpurun_out
STAGES="c6b rcab1 rcab2 rcab6 ca6 cab1 cab6 long edsr rcan"
for s in $STAGES; do
  echo "=== $s"
  timeout 200 python scripts/cluster_debug.py $s 2>&1 | tail -16
done 2>&1 | tee gpurun_out/cluster_debug.txt
(timeout 200 python scripts/cluster_trace.py group) 2>&1 | tee gpurun_out/cluster_trace.txt
timeout 300 python bench.py --workload train --steps 20 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_train.log
