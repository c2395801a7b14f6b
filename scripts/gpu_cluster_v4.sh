mkdir -p gpurun_out
(cd scripts/probes && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../sr-pytorch-lightning_b200/csrc probes.cu -o /tmp/probes_bin)
timeout 60 /tmp/probes_bin p10 2>&1 | grep P10 | tee gpurun_out/r02_hw_probes_p10.txt
STAGES="c1x3 c2h c2v c4 c6b c8 c2h c6b c8 c2h c6b c8 long ca1 ca6 cab1 cab6 cab6 cab6 edsr rcan"
for s in $STAGES; do
  echo "=== $s"
  timeout 150 python scripts/cluster_debug.py $s 2>&1 | tail -14
done 2>&1 | tee gpurun_out/cluster_debug.txt
(timeout 200 python scripts/cluster_trace.py long; timeout 200 python scripts/cluster_trace.py group) 2>&1 | tee gpurun_out/cluster_trace.txt
