mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-sweep > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_r01c.err
python tools/profile_step.py > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r01c.csv python tools/profile_step.py > gpurun_out/ncu1.log 2>&1; echo "launch-list rc=$?"
