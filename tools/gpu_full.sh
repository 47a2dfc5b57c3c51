mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc=$?"
tail -6 gpurun_out/t_all.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r01d.err
