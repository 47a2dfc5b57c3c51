mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -x -p no:cacheprovider > gpurun_out/t_full.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_full.log
