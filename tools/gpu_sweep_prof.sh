mkdir -p gpurun_out
timeout 600 python tools/profile_sweep.py > gpurun_out/sweep_prof.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/sweep_prof.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; echo "bench rc=$?"; cat gpurun_out/bench_tmp.json
