mkdir -p gpurun_out
N=${1:-2}
export TORCH_NCCL_ASYNC_ERROR_HANDLING=0
( timeout 600 python -m pytest tests/test_tiled_gpu.py -x -q 2>&1 | tail -4 ) | tee gpurun_out/peer3_tests.log
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/tiled_peer_check.py large 30 2>&1 | grep -E "PEER|MISMATCH|Error" | head ) | tee gpurun_out/peer3_large_$N.log
