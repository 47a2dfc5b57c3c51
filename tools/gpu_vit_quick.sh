mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vit.py tests/test_gpu_ops.py -q -x -p no:cacheprovider > gpurun_out/t_vit.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_vit.log
timeout 300 python tools/bench_vit.py --batch 256 --steps 5 --warmup 3 > gpurun_out/vit_n1.json 2> gpurun_out/vit_n1.err; cat gpurun_out/vit_n1.json; tail -3 gpurun_out/vit_n1.err
