mkdir -p gpurun_out
N=${1:-2}
export TORCH_NCCL_ASYNC_ERROR_HANDLING=0
run() { name=$1; shift; ( env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/dp_${name}_$N.json; python - "$name" gpurun_out/dp_${name}_$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    i=d.get("infer4k") or {}
    print(sys.argv[1], "train", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "| infer4k", round(i.get("value",0),2), "fps e2e", round((i.get("e2e") or {}).get("value",0),2), "launches", i.get("gpu_launches"))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[2]).read()[-600:])
PY
}
run buckets2 A=1
run nobuckets SRB200_ALLREDUCE_BUCKETS=0
