# N = 128 weight-gradient kernel: parity tests, then the RCAN / EDSR steps and the group launch timed alone under ncu
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
run rcan_a X=1
MODEL=edsr run edsr X=1
run rcan_b X=1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -k regex:wgrad_umma -s 2 -c 1 python scripts/ncu_targets.py wgrad 2>&1 | grep -E "gpu__time|cycles_elapsed"
