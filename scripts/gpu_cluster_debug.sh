# per-sample cluster chain kernel: stage-by-stage parity against the per-layer kernels (one process per stage)
mkdir -p gpurun_out
STAGES="${STAGES:-c1 c1x3 c2h c2v c4 c6 c6b c8 long ca1 ca6 cab1 cab6 edsr rcan}"
for s in $STAGES; do
  echo "=== $s"
  timeout 150 python scripts/cluster_debug.py $s 2>&1 | tail -25
done 2>&1 | tee gpurun_out/cluster_debug.txt
