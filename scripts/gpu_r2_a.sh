# round 2, call A: DSMEM (P7) + cta_group::2 (P8) probes, the stock PyTorch/cuDNN "library" bar, baseline bench line
set -x
mkdir -p gpurun_out
(cd scripts/probes && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../sr-pytorch-lightning_b200/csrc probes.cu -o /tmp/probes_bin)
timeout 60 /tmp/probes_bin p7 2>&1 | tee gpurun_out/r02_hw_probes_p7.txt
timeout 60 /tmp/probes_bin p8 2>&1 | tee gpurun_out/r02_hw_probes_p8.txt
timeout 300 python bench.py --impl library --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/r02_bench_library_rcan.json
timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02_bench_rcan_v0.json
