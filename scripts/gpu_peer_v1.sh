mkdir -p gpurun_out
N=${1:-2}
export TORCH_NCCL_ASYNC_ERROR_HANDLING=0
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/tiled_peer_check.py small 20 2>&1 | tail -12 ) | tee gpurun_out/peer_small_$N.log
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/tiled_peer_check.py large 30 2>&1 | tail -12 ) | tee gpurun_out/peer_large_$N.log
# data-parallel training with bucketed all-reduce: same data on both ranks -> losses must match the 1-GPU trace
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/overlap_determinism.py 12 2>&1 | grep -E "losses|grad norm|Error|error" | head -8 ) | tee gpurun_out/dp_losses_$N.log
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --workload train --steps 20 --warmup 5 --no-extras 2>&1 | tail -1 | cut -c1-600 ) | tee gpurun_out/dp_bench_$N.log
( SRB200_ALLREDUCE_BUCKETS=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --workload train --steps 20 --warmup 5 --no-extras 2>&1 | tail -1 | cut -c1-400 ) | tee gpurun_out/dp_bench_nobuckets_$N.log
