# 8-GPU: strip-tiled 4K inference with halo exchange, DDP training bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_infer4k.py --steps 5 --warmup 2 > gpurun_out/infer4k_8gpu.json 2> gpurun_out/infer4k_8gpu.err; echo "rc=$?"; tail -n 2 gpurun_out/infer4k_8gpu.err | cut -c1-300; cat gpurun_out/infer4k_8gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_rcan_8gpu.json 2> gpurun_out/bench_rcan_8gpu.err; echo "rc=$?"; tail -n 2 gpurun_out/bench_rcan_8gpu.err | cut -c1-300; cut -c1-400 gpurun_out/bench_rcan_8gpu.json
