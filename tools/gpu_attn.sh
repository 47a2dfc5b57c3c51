mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_ops.py -q -x -k "attention" -p no:cacheprovider > gpurun_out/t_attn.log 2>&1; echo "attn rc=$?"
tail -15 gpurun_out/t_attn.log
timeout 120 python tools/bench_attn.py > gpurun_out/bench_attn.json 2> gpurun_out/bench_attn.err; echo "bench rc=$?"; cat gpurun_out/bench_attn.json; tail -3 gpurun_out/bench_attn.err
