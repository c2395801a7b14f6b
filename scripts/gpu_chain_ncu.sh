set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_chain -s 2 -c 1 -o gpurun_out/prof_chain python scripts/chain_bench.py > gpurun_out/ncu_chain.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/ncu_chain.log
ls -la gpurun_out/*.ncu-rep
