mkdir -p gpurun_out
N=${1:-8}
export TORCH_NCCL_ASYNC_ERROR_HANDLING=0
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/tiled_peer_check.py large 40 2>&1 | grep -E "PEER|MISMATCH|Error" | head ) | tee gpurun_out/peer4_large_$N.log
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline --workload infer4k 2>&1 | tail -1 ) > gpurun_out/scale_infer_peer2_$N.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/scale_infer_peer2_*.json")):
    try:
        d=json.loads(open(f).read()); i=d["infer4k"]
        print(f, round(i["value"],2), "fps e2e", round(i["e2e"]["value"],2), "launches", i["gpu_launches"], i["clocks"])
    except Exception as e:
        print(f, "FAILED", e, open(f).read()[-500:])
PY
