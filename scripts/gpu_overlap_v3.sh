mkdir -p gpurun_out
summ='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "fwd", round(r["us_forward_launch"],1), "bwd", round(r["us_backward_launch"],1), "loss", d["config"]["loss_last"])'
run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline --no-extras --sustain-seconds 0.5 2>/dev/null | python -c "$summ" "$name" | tee -a gpurun_out/overlap_bench.log; }
rm -f gpurun_out/overlap_bench.log
run default A=1
run no_overlap SRB200_WGRAD_OVERLAP=0
run fwd_cluster SRB200_CHAIN_FWD=cluster
run sms52_g7 SRB200_WGRAD_OVERLAP_GROUPS=7
run sms52_g9 SRB200_WGRAD_OVERLAP_GROUPS=9
run sms48_g8 SRB200_WGRAD_OVERLAP_SMS=48
run sms50_g8 SRB200_WGRAD_OVERLAP_SMS=50
