mkdir -p gpurun_out
CMD="python tools/bench_vit.py --batch 256 --steps 1 --warmup 3 --profile --no-graph"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vit_pipe.csv $CMD > gpurun_out/ncu_a.log 2>&1; echo "A rc=$?"
HBA_GEMM_DEBUG=4 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vit_nopipe.csv $CMD > gpurun_out/ncu_b.log 2>&1; echo "B rc=$?"
STEP="python tools/profile_step.py"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_pipe.csv $STEP > gpurun_out/ncu_c.log 2>&1; echo "C rc=$?"
HBA_GEMM_DEBUG=4 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_nopipe.csv $STEP > gpurun_out/ncu_d.log 2>&1; echo "D rc=$?"
