mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_model.py -q -x -p no:cacheprovider > gpurun_out/t_pipe.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_pipe.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; echo "bench rc=$?"; cat gpurun_out/bench_r01e.json; tail -3 gpurun_out/bench_r01e.err
