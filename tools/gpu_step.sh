mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc=$?"
tail -4 gpurun_out/t_all.log
python bench.py --steps 20 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; echo "bench rc=$?"; cat gpurun_out/bench_tmp.json
python tools/profile_step.py > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_tmp.csv python tools/profile_step.py > gpurun_out/ncu1.log 2>&1; echo "launch-list rc=$?"
