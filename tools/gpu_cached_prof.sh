mkdir -p gpurun_out
python tools/profile_cached_step.py > gpurun_out/plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cached_r01e.csv python tools/profile_cached_step.py > gpurun_out/ncu1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/plain.log
timeout 300 python bench.py --sweep-only > gpurun_out/sweep_only.json 2> gpurun_out/sweep_only.err; echo "rc=$?"; cat gpurun_out/sweep_only.json
