# N = 128 weight-gradient kernel: parity tests, then A/B of the RCAN / EDSR steps
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "wgrad" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_models_gpu.py tests/test_trainer_gpu.py -x -q 2>&1 | tail -3
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --workload train --no-extras --no-cpu-baseline --sustain-seconds 0.5 --model ${MODEL:-rcan} 2>&1 | tail -1 > gpurun_out/wn_$label.json
  python -c "
import json
d=json.load(open('gpurun_out/wn_$label.json')); print('$label', round(d['value'],1), round(d['ms_per_step'],3))"
}
run scatter_g8 SRB200_WGRAD_SCATTER=1
run staged_g8 SRB200_WGRAD_SCATTER=0
run staged_g9 SRB200_WGRAD_OVERLAP_GROUPS=9
run staged_g10 SRB200_WGRAD_OVERLAP_GROUPS=10
run staged_g10_sm44 SRB200_WGRAD_OVERLAP_GROUPS=10 SRB200_WGRAD_OVERLAP_SMS=44
MODEL=edsr run edsr_staged SRB200_WGRAD_SCATTER=0
MODEL=rdn run rdn SRB200_WGRAD_SCATTER=0
