mkdir -p gpurun_out
for v in "A=1" "A=2" "SRB200_WGRAD_OVERLAP=0" "SRB200_WGRAD_OVERLAP=0" "SRB200_CHAIN_CLUSTER=0" "SRB200_CHAIN_FWD=cluster" "SRB200_CHAIN_FWD=cluster SRB200_WGRAD_OVERLAP=0"; do
  echo "== $v"; env $v timeout 200 python scripts/overlap_determinism.py 12 2>&1 | tail -2
done | tee gpurun_out/overlap_determinism.log
