mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_tests.log 2>&1
tail -4 gpurun_out/final_tests.log
( time timeout 900 python bench.py ) > gpurun_out/final_bench_default.log 2>&1
tail -c 400 gpurun_out/final_bench_default.log
# launch list of one captured training step (device time per kernel; shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 1500 --csv --log-file gpurun_out/r02_launches_rcan.csv python bench.py --workload train --steps 1 --warmup 1 --no-extras --no-cpu-baseline --sustain-seconds 0 > gpurun_out/final_ncu_launches.log 2>&1
tail -2 gpurun_out/final_ncu_launches.log | cut -c1-200
# full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_cluster -s 2 -c 1 -f -o gpurun_out/r02_chain_cluster_bwd python scripts/ncu_targets.py group 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_chain_kernel -s 2 -c 1 -f -o gpurun_out/r02_chain_flags_fwd python scripts/ncu_targets.py group 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wide -s 2 -c 1 -f -o gpurun_out/r02_conv_wide128 python scripts/ncu_targets.py wide128 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_umma -s 2 -c 2 -f -o gpurun_out/r02_wgrad_group python scripts/ncu_targets.py wgrad 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_umma -s 4 -c 2 -f -o gpurun_out/r02_wgrad_group_52sm python scripts/ncu_targets.py wgrad 52 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
