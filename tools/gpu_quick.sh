mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "dora or attention or layernorm" -p no:cacheprovider > gpurun_out/t_q.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_q.log
python tools/bench_kernels.py --reps 20 > gpurun_out/kernels_r01.json 2> gpurun_out/kernels_r01.err; tail -3 gpurun_out/kernels_r01.err
