mkdir -p gpurun_out
# ncu launch lists (gpu__time_duration.sum, --clock-control none) of the final tree
BENCH="python bench.py --steps 2 --warmup 3 --no-sweep --no-vit --no-cpu-baseline"
$BENCH > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r01e.csv $BENCH > gpurun_out/ncu_bench.log 2>&1; echo "bench launch list rc=$?"
STEP="python tools/profile_step.py"
$STEP > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_r01e.csv $STEP > gpurun_out/ncu1.log 2>&1; echo "step launch list rc=$?"
CACHED="python tools/profile_cached_step.py"
$CACHED > gpurun_out/plain_c.log 2>&1 && timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cached_r01e.csv $CACHED > gpurun_out/ncu_c.log 2>&1; echo "cached launch list rc=$?"
VIT="python tools/bench_vit.py --batch 256 --steps 1 --warmup 3 --profile --no-graph"
$VIT > gpurun_out/plain_vit.log 2>&1 && timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vit_r01e.csv $VIT > gpurun_out/ncu_vit.log 2>&1; echo "vit launch list rc=$?"
du -sh gpurun_out
