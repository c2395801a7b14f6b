mkdir -p gpurun_out
STAGES="rcab6 cab1 cab6 rcan"
for s in $STAGES; do
  echo "=== $s"
  timeout 200 python scripts/cluster_debug.py $s 2>&1 | tail -16
done 2>&1 | tee gpurun_out/cluster_debug.txt
(timeout 200 python scripts/cluster_trace.py group) 2>&1 | tee gpurun_out/cluster_trace.txt
timeout 300 python bench.py --workload train --steps 20 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], r['us_forward_launch'], r['us_backward_launch'])" | tee gpurun_out/bench_train.log
