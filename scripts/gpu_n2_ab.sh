export TORCH_NCCL_ASYNC_ERROR_HANDLING=0
run() { name=$1; shift; ( env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-cpu-baseline --workload train 2>&1 | tail -1 ) | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"; }
run default A=1
run nocoop SRB200_CHAIN_COOP=0
run default2 A=1
run nocoop2 SRB200_CHAIN_COOP=0
run adopt SRB200_WGRAD_ADOPT=1
