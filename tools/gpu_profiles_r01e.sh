mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-sweep --no-vit --no-cpu-baseline"
$BENCH > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r01e.csv $BENCH > gpurun_out/ncu_bench.log 2>&1; echo "bench launch list rc=$?"
STEP="python tools/profile_step.py"
$STEP > gpurun_out/plain.log 2>&1 || exit 1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_r01e.csv $STEP > gpurun_out/ncu1.log 2>&1; echo "step launch list rc=$?"
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -o gpurun_out/gemm_step_r01e -f $STEP > gpurun_out/ncu_gemm_full.log 2>&1; echo "gemm full rc=$?"; tail -2 gpurun_out/ncu_gemm_full.log
ls -la gpurun_out/gemm_step_r01e.ncu-rep
