# A/B: non-trunk weight gradients adopted onto the side stream under the first backward chain (SRB200_WGRAD_ADOPT)
mkdir -p gpurun_out
run() {
  label=$1; shift
  env "$@" timeout 200 python bench.py --workload train --no-extras --no-cpu-baseline --sustain-seconds 0.3 2>&1 | tail -1 > gpurun_out/ad_$label.json
  python -c "
import json
d=json.load(open('gpurun_out/ad_$label.json')); print('$label', round(d['value'],1), round(d['ms_per_step'],4))"
}
run off SRB200_WGRAD_ADOPT=0
run on SRB200_WGRAD_ADOPT=1
run on_g10 SRB200_WGRAD_ADOPT=1 SRB200_WGRAD_OVERLAP_GROUPS=10
SRB200_WGRAD_ADOPT=1 timeout 200 python -m pytest tests/test_trainer_gpu.py tests/test_runner_gpu.py -x -q 2>&1 | tail -2
