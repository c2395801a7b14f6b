mkdir -p gpurun_out
N=${1:-8}
export TORCH_NCCL_ASYNC_ERROR_HANDLING=0
run() { name=$1; wl=$2; shift; shift; ( env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline --workload $wl 2>&1 | tail -1 ) > gpurun_out/scale_${name}_$N.json; python - "$name" gpurun_out/scale_${name}_$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    i=d.get("infer4k") or {}
    print(sys.argv[1], "train", round(d.get("value") or 0,1), "ms", round(d.get("ms_per_step") or 0,3), "e2e", round((d.get("e2e") or {}).get("value",0),1), "| infer4k", round(i.get("value",0),2), "fps e2e", round((i.get("e2e") or {}).get("value",0),2), "clocks", (i.get("clocks") or d.get("clocks") or {}).get("sm_mhz"))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[2]).read()[-800:])
PY
}
run train_default train A=1
run train_buckets2 train SRB200_ALLREDUCE_BUCKETS=1
run infer_peer infer4k A=1
run infer_nccl infer4k SRB200_TILED_EXCHANGE=nccl
