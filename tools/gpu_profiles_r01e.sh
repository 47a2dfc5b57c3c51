mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-sweep --no-vit --no-cpu-baseline"
$BENCH > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r01e.csv $BENCH > gpurun_out/ncu_bench.log 2>&1; echo "bench launch list rc=$?"
STEP="python tools/profile_step.py"
$STEP > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_r01e.csv $STEP > gpurun_out/ncu1.log 2>&1; echo "step launch list rc=$?"
# one vision block's GEMMs (qkv, out_proj, c_fc, c_proj) x 2 blocks, full set; the report stays on the box
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 41 -c 8 -o /tmp/gemm_step_r01e -f $STEP > gpurun_out/ncu_gemm_full.log 2>&1; echo "gemm full rc=$?"
ncu -i /tmp/gemm_step_r01e.ncu-rep --page raw --csv > gpurun_out/gemm_step_r01e_raw.csv 2>/dev/null; ls -la /tmp/gemm_step_r01e.ncu-rep gpurun_out/gemm_step_r01e_raw.csv
VIT="python tools/bench_vit.py --batch 256 --steps 1 --warmup 3 --profile --no-graph"
$VIT > gpurun_out/plain_vit.log 2>&1 && timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vit_r01e.csv $VIT > gpurun_out/ncu_vit.log 2>&1; echo "vit launch list rc=$?"
du -sh gpurun_out
