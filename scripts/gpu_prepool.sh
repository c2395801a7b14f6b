mkdir -p gpurun_out
for s in rcabs rcab1 rcabr rcab6; do
  echo "=== $s (flags kernel, prepool)"
  SRB200_CHAIN_CLUSTER=0 timeout 200 python scripts/cluster_debug.py $s 2>&1 | tail -7
done 2>&1 | tee gpurun_out/prepool_debug.txt
summ='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "fwd", round(r["us_forward_launch"],1), "bwd", round(r["us_backward_launch"],1), "loss", d["config"]["loss_last"])'
run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline --no-extras --sustain-seconds 0.5 2>/dev/null | python -c "$summ" "$name" | tee -a gpurun_out/prepool_bench.log; }
rm -f gpurun_out/prepool_bench.log
run prepool A=1
run no_prepool SRB200_CHAIN_CA_PREPOOL=0
