# round 2, first thing: DSMEM latency probes for the per-sample cluster design; not run in round 1
set -x
mkdir -p gpurun_out
cd scripts/probes && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../sr-pytorch-lightning_b200/csrc probes.cu -o /tmp/probes_bin && cd ../..
timeout 60 /tmp/probes_bin p7 | tee gpurun_out/hw_probes_p7.txt
