# the driver's multi-GPU launch of bench.py (default workload: training line + infer4k key) on N GPUs of one box
mkdir -p gpurun_out
N=${1:-2}
export TORCH_NCCL_ASYNC_ERROR_HANDLING=0
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/scale_default_$N.log 2>&1
tail -1 gpurun_out/scale_default_$N.log > /dev/null
python - $N <<'PY'
import json,sys
n=sys.argv[1]
for l in open(f"gpurun_out/scale_default_{n}.log"):
    if l.startswith("{"):
        d=json.loads(l); i=d.get("infer4k") or {}
        print("N", n, "train", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "gpu_data", round((d.get("e2e_gpu_data") or {}).get("value",0),1),
              "| infer4k", round(i.get("value",0),2), "e2e", round((i.get("e2e") or {}).get("value",0),2), "clocks", (d.get("clocks") or {}).get("sm_mhz"), (i.get("clocks") or {}).get("sm_mhz"))
PY
grep -E "^real" gpurun_out/scale_default_$N.log
