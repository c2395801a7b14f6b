# full validation on one B200: GPU tests, the default bench line (both timed), optionally the ncu launch list and full captures
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_tests.log 2>&1
grep -E "passed|failed|error" gpurun_out/final_tests.log | tail -3
( time timeout 900 python bench.py ) > gpurun_out/final_bench_default.log 2>&1
tail -c 300 gpurun_out/final_bench_default.log
if [ "$1" = "ncu" ]; then
  # launch list of the timed step only (device time per kernel; serialised: use for shares)
  SRB200_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_rcan.csv python bench.py --workload train --steps 1 --warmup 1 --no-extras --no-cpu-baseline --sustain-seconds 0 > gpurun_out/final_ncu_launches.log 2>&1
  for t in "chain_cluster group r02_chain_cluster_bwd" "conv_chain_kernel group r02_chain_flags_fwd" "conv_wide wide128 r02_conv_wide128" "wgrad_umma wgrad r02_wgrad_group"; do
    set -- $t
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -f -o gpurun_out/$3 python scripts/ncu_targets.py $2 2>&1 | tail -1
  done
fi
