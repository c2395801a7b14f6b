mkdir -p gpurun_out
STAGES="c1x3 c2h c2v c4 c6b c8 c6b c8 long ca1 ca6 cab1 cab6 cab6 edsr rcan"
for s in $STAGES; do
  echo "=== $s"
  timeout 150 python scripts/cluster_debug.py $s 2>&1 | tail -14
done 2>&1 | tee gpurun_out/cluster_debug.txt
(timeout 200 python scripts/cluster_trace.py long; timeout 200 python scripts/cluster_trace.py group) 2>&1 | tee gpurun_out/cluster_trace.txt
