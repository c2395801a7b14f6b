mkdir -p gpurun_out
for v in relu res mask; do timeout 120 python scripts/trace_c64.py $v; done > gpurun_out/trace_c64.txt 2>&1
SRB200_NO_PDL=1 timeout 120 python scripts/trace_c64.py relu >> gpurun_out/trace_c64.txt 2>&1
cat gpurun_out/trace_c64.txt
