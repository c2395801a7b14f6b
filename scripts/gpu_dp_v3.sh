mkdir -p gpurun_out
N=${1:-2}
export TORCH_NCCL_ASYNC_ERROR_HANDLING=0
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/overlap_determinism.py 12 2>&1 | grep -E "losses|grad norm|Error|error" | head -8 ) | tee gpurun_out/dp3_losses_$N.log
run() { name=$1; shift; ( env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline --workload train 2>&1 | tail -1 ) > gpurun_out/dp3_${name}_$N.json; python - "$name" gpurun_out/dp3_${name}_$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    print(sys.argv[1], "train", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "loss", d["config"]["loss_last"])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[2]).read()[-800:])
PY
}
run tail A=1
run off SRB200_ALLREDUCE_BUCKETS=0
run tail8 SRB200_ALLREDUCE_SMS=8
