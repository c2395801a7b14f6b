# 2-GPU validation at HEAD: strip-tiled 4K inference with halo exchange, DDP training bench (NCCL all-reduce in the graph), tiled tests
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_infer4k.py --steps 5 --warmup 2 > gpurun_out/infer4k_2gpu.json 2> gpurun_out/infer4k_2gpu.err; echo "rc=$?"; tail -n 3 gpurun_out/infer4k_2gpu.err; cat gpurun_out/infer4k_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_rcan_2gpu.json 2> gpurun_out/bench_rcan_2gpu.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_rcan_2gpu.err; cut -c1-400 gpurun_out/bench_rcan_2gpu.json
timeout 300 python -m pytest tests/test_tiled_gpu.py tests/test_kernels_gpu.py -q -m gpu -k "tiled or bias_grads or relu" --tb=short -p no:cacheprovider 2>&1 | tail -n 4
