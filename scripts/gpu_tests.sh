set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --tb=short -k "not umma and not dgrad and not deferred" -p no:cacheprovider > gpurun_out/t1_simt.log 2>&1; echo "rc=$?" >> gpurun_out/t1_simt.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --tb=short -k "umma or dgrad or deferred" -p no:cacheprovider > gpurun_out/t2_umma.log 2>&1; echo "rc=$?" >> gpurun_out/t2_umma.log
timeout 900 python -m pytest tests/test_models_gpu.py -q --tb=short -k "fp32" -p no:cacheprovider > gpurun_out/t3_models_fp32.log 2>&1; echo "rc=$?" >> gpurun_out/t3_models_fp32.log
timeout 900 python -m pytest tests/test_models_gpu.py -q --tb=short -k "not fp32" -p no:cacheprovider > gpurun_out/t4_models_bf16.log 2>&1; echo "rc=$?" >> gpurun_out/t4_models_bf16.log
tail -5 gpurun_out/t1_simt.log gpurun_out/t2_umma.log gpurun_out/t3_models_fp32.log gpurun_out/t4_models_bf16.log
