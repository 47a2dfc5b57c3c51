mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "not gemm" -p no:cacheprovider > gpurun_out/t_ops.log 2>&1; echo "ops rc=$?" >> gpurun_out/rc.txt
timeout 240 python -m pytest tests/test_gpu_ops.py -q -k "gemm" -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?" >> gpurun_out/rc.txt
timeout 300 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider > gpurun_out/t_model.log 2>&1; echo "model rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
