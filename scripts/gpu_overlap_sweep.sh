# weight-gradient overlap: trainer tests, bench variants (one line each), CUPTI timeline of the default configuration
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_runner_gpu.py -x -q 2>&1 | tail -3 ) | tee gpurun_out/overlap_tests.log
summ='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "fwd", round(r["us_forward_launch"],1), "bwd", round(r["us_backward_launch"],1), "loss", d["config"]["loss_last"])'
run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline --no-extras --sustain-seconds 0.5 2>/dev/null | python -c "$summ" "$name" | tee -a gpurun_out/overlap_bench.log; }
rm -f gpurun_out/overlap_bench.log
run default A=1
for v in "$@"; do run "$v" $v; done
timeout 300 python scripts/step_timeline.py gpurun_out/step_timeline.txt > /dev/null 2>&1
